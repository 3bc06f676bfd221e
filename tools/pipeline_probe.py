"""GPU probe: the phase pipeline (BLCD_PIPELINE=1) against the fused kernel -- equality of results and time per env step.
    python tools/pipeline_probe.py [env] [worlds] [T]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import boxlcd_b200 as blcd
from boxlcd_b200.vec_env import VecWorldEnv

name = sys.argv[1] if len(sys.argv) > 1 else 'Urchin'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20


def make(pipeline, n, **kw):
  os.environ['BLCD_PIPELINE'] = '1' if pipeline else '0'
  v = VecWorldEnv(blcd.env_map[name](), n, seed=0, **kw)
  os.environ.pop('BLCD_PIPELINE')
  return v


def ev(fn):
  torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record(); fn(); b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)


# equality on a small batch
m = 8192
f, p = make(False, m), make(True, m)
f.reset_dev(); p.reset_dev()
rf, rp = f.rollout_dev(12), p.rollout_dev(12)
torch.cuda.synchronize()
for k in rf:
  same = (rf[k] == rp[k]).reshape(m, -1).all(1).float().mean().item()
  print(f'{name} rollout {k}: worlds identical {same:.5f}', flush=True)
print('bodies identical:', float((f.get_bodies() == p.get_bodies()).all(2).all(1).mean()), 'counters identical:', float((f.counters() == p.counters()).all(1).mean()), flush=True)
# stepping API
act = torch.rand((m, f.A), device='cuda') * 2 - 1
of, _ = f.step_dev(act, observe=True); of = {k: v.clone() for k, v in of.items()}
op, _ = p.step_dev(act, observe=True)
print('step_observe identical:', {k: bool((of[k] == op[k]).all()) for k in of}, flush=True)
del f, p
for pipeline in (False, True):
  v = make(pipeline, n)
  v.reset_dev(); v.rollout_dev(3)
  for rep in range(2):
    v.reset_dev()
    ms = ev(lambda: v.rollout_dev(T))
    print(f'{name} n={n} pipeline={pipeline}: {ms / T:.3f} ms/env-step, {n * T / ms / 1e3:.2f} M env-steps/s', flush=True)
  del v
