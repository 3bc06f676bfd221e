"""GPU probe behind tests/test_gpu_parity.py::test_single_env_step_within_tolerance_of_oracle: for every world of the
single-step protocol dump the position / angle error against the oracle together with the diagnostic counters of both
sides (contacts, position iterations, TOI events / calls, manifold points), so that the worlds above the 1e-4 bar can be
classified (decision flip vs round-off).  Run once per library build:
    python tools/parity_tail_probe.py out.npz            (BLCD_LIB=... selects an experiment build, e.g. FMAD=false)
"""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from common import make_env, random_bodies, rel_err, ENVS_CORE  # noqa: E402
from oracle import oracle  # noqa: E402
from boxlcd_b200.vec_env import VecWorldEnv  # noqa: E402


def main(path, n=4096):
  out = {}
  for name in ENVS_CORE + ['CrabCube', 'SpiderCube']:
    env = make_env(name)
    rng = np.random.RandomState(3)
    bodies, variants = random_bodies(env, n, rng)
    act = rng.uniform(-1.2, 1.2, (n, env.act_size)).astype(np.float32)
    ow = oracle.OracleWorlds(env.layout.spec, n, threads=os.cpu_count() or 1)
    ow.set_bodies(bodies, variants)
    ow.step(act)
    ref = ow.get_bodies()
    v = VecWorldEnv(env, n)
    v.set_bodies(bodies, variants)
    v.step_dev(torch.as_tensor(act).cuda(), observe=False)
    got = v.get_bodies()
    out[f'{name}_ref'], out[f'{name}_got'] = ref, got
    out[f'{name}_cnt_ref'], out[f'{name}_cnt_got'] = ow.counters(), v.counters()
    rel = rel_err(got[..., :3], ref[..., :3]).max((1, 2))
    flip = (out[f'{name}_cnt_ref'] != out[f'{name}_cnt_got']).any(1)
    print(f'{name}: >1e-4 rel: {(rel > 1e-4).sum()} of {n}; counters differ: {flip.sum()}; >1e-4 with equal counters: {((rel > 1e-4) & ~flip).sum()}; '
          f'worst non-flipped {rel[~flip].max():.2e}; worst flipped abs {np.abs(got[..., :3] - ref[..., :3]).max((1, 2))[flip].max() if flip.any() else 0:.2e}', flush=True)
  np.savez_compressed(path, **out)


if __name__ == '__main__':
  main(sys.argv[1])
