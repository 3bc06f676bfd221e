"""Device-resident random-action data collection with the reference's dataset formats.

  python -m boxlcd_b200.collect --env=Urchin --collect_n=10000           -> rollouts/Urchin-10000.npz
  python -m boxlcd_b200.collect --env=Urchin --barrels=10 --logdir=logs/  -> logs/{train}/{timestamp}-{ep_len}.barrel.npz

Replaces the loops of examples/collect.py:24-41 and research/data.py:36-79: every rollout is one world of a batched
fused kernel launch (blcd_rollout: on-device action RNG -> step -> obs -> frame); the host only reshapes and writes
(npz_writer.py: the same .npz container, deflated on all host cores instead of one).
File formats are the reference's: `action` float64 [N, T, A], `full_state` float32 [N, T, S], `proprio` float32
[N, T, P], `lcd` bool [N, T, H, W]; element [i, j] is the observation BEFORE action j; no resets inside a rollout.
With several GPUs (torchrun) each rank collects a contiguous shard of rollouts; npz mode gathers the shards on rank 0
(the only collective on this path), barrel mode writes per-rank barrel files (already a sharded format).
"""
import argparse
import os
import pathlib
import sys
import time
from datetime import datetime

import numpy as np
from boxlcd_b200.npz_writer import savez_compressed_parallel

BARREL_SIZE = 1000  # research/data.py:21


def config():
  """the collect-relevant subset of examples/utils.py:13-37 plus the env defaults"""
  from boxlcd_b200 import ENV_DG
  from boxlcd_b200.utils import AttrDict
  G = AttrDict()
  G.logdir = pathlib.Path('./logs/')
  G.datadir = pathlib.Path('.')
  G.collect_n = 10000
  G.env = 'Bounce'
  G.barrels = 0          # > 0: write this many 1000-rollout barrel files under logdir/<split>/ instead of one npz
  G.split = 'train'
  G.batch = 65536        # worlds per launch
  G.seed = 0
  for key, val in ENV_DG.items():
    assert key not in G, key
    G[key] = val
  return G


def parse_args(argv=None):
  """examples/utils.py:39-47: flags from the config dict, then the chosen env's ENV_DG as defaults, then re-parse"""
  from boxlcd_b200 import env_map
  from boxlcd_b200.utils import AttrDict, args_type
  parser = argparse.ArgumentParser()
  for key, value in config().items():
    parser.add_argument(f'--{key}', type=args_type(value), default=value)
  temp = parser.parse_args(argv)
  parser.set_defaults(**env_map[temp.env].ENV_DG)
  return AttrDict(parser.parse_args(argv).__dict__)


def shard_range(n_items, rank, world):
  """contiguous [lo, hi) of n_items owned by `rank` of `world` (rollouts are keyed by their global index, so the
  dataset does not depend on the number of ranks)"""
  return n_items * rank // world, n_items * (rank + 1) // world


def gather_shards(data, rank, world):
  """final gather of per-rank dataset shards on rank 0 -- the only collective on this path"""
  if world == 1:
    return data
  import torch.distributed as dist
  parts = [None] * world if rank == 0 else None
  dist.gather_object(data, parts, dst=0)
  if rank != 0:
    return None
  return {k: np.concatenate([p[k] for p in parts]) for k in data}


def collect_arrays(env, n_rollouts, T, batch=65536, seed=0, device=None, world_offset=0, progress=None):
  """n_rollouts x T random-action rollouts of `env` -> dict of host arrays in the reference layout."""
  import torch
  from boxlcd_b200.vec_env import VecWorldEnv
  S, A, P = env.obs_size, env.act_size, max(env.pobs_size, 1)
  H, W = env.observation_space.spaces['lcd'].shape
  # np.empty: every element is written below; the copies land straight in these arrays (no intermediate host tensors)
  out = {'action': np.empty((n_rollouts, T, A), np.float64), 'full_state': np.empty((n_rollouts, T, S), np.float32),
         'proprio': np.empty((n_rollouts, T, P), np.float32), 'lcd': np.empty((n_rollouts, T, H, W), np.bool_)}
  pidx = torch.as_tensor(np.asarray(env.pobs_idxs, np.int64)) if env.pobs_size else None
  done, t0 = 0, time.time()
  vec = None
  while done < n_rollouts:
    n = min(batch, n_rollouts - done)
    if vec is None or vec.n != n:
      if vec is not None:
        vec.close()
      vec = VecWorldEnv(env, n, device=device, seed=seed, world_offset=world_offset + done)
    else:   # same batch size: keep the allocation, move the RNG window
      vec.close()
      vec = VecWorldEnv(env, n, device=device, seed=seed, world_offset=world_offset + done)
    vec.reset_dev()
    r = vec.rollout_dev(T)
    # conversions (f32 -> f64 actions, proprio gather, bit unpacking) run on the GPU; each result is copied directly into
    # its slice of the output array
    sl = slice(done, done + n)
    torch.from_numpy(out['action'][sl]).copy_(r['action'].double())
    torch.from_numpy(out['full_state'][sl]).copy_(r['full_state'])
    if pidx is not None:
      torch.from_numpy(out['proprio'][sl]).copy_(r['full_state'].index_select(2, pidx.to(r['full_state'].device)))
    else:
      out['proprio'][sl] = 0.0
    torch.from_numpy(out['lcd'][sl]).copy_(vec.unpack_lcd(r['lcd_bits']))
    done += n
    if progress:
      progress(done, n_rollouts, done * T / (time.time() - t0))
  if vec is not None:
    vec.close()
  return out


def main(argv=None):
  import torch
  import boxlcd_b200 as blcd
  G = parse_args(argv)
  rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
  if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    dist.init_process_group('nccl')
  env = blcd.env_map[G.env](G)
  T = G.ep_len
  say = (lambda d, n, fps: print(f'\r{d}/{n} rollouts, {fps:,.0f} env-steps/s', end='', flush=True)) if rank == 0 else None
  if G.barrels > 0:
    logdir = pathlib.Path(G.logdir) / G.split
    logdir.mkdir(parents=True, exist_ok=True)
    for bi in range(rank, G.barrels, world):
      data = collect_arrays(env, BARREL_SIZE, T, min(G.batch, BARREL_SIZE), G.seed, world_offset=bi * BARREL_SIZE, progress=say)
      stamp = datetime.now().strftime('%Y%m%dT%H%M%S') + (f'r{rank}b{bi}' if world > 1 or G.barrels > 1 else '')
      savez_compressed_parallel(logdir / f'{stamp}-{T}.barrel', **data)
  else:
    N = G.collect_n
    lo, hi = shard_range(N, rank, world)
    data = collect_arrays(env, hi - lo, T, G.batch, G.seed, world_offset=lo, progress=say)
    data = gather_shards(data, rank, world)
    if rank == 0:
      os.makedirs('rollouts', exist_ok=True)
      savez_compressed_parallel(f'rollouts/{G.env}-{N}.npz', **data)
      print(f'\nwrote rollouts/{G.env}-{N}.npz')
  if world > 1:
    import torch.distributed as dist
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
