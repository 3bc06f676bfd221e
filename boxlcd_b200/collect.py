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


def gather_shards(data, rank, world, n_total=None, chunk_bytes=256 << 20):
  """Final gather of the per-rank dataset shards on rank 0 -- the only communication on this path.  Shards are contiguous
  rollout ranges (shard_range), so rank 0 allocates the full arrays once and every other rank streams its arrays into
  their slice with point-to-point sends of at most `chunk_bytes` (NCCL over NVLink: device tensors staged through one
  reusable buffer per side; gloo: host tensors) -- no pickling, no second copy of a multi-GB array."""
  if world == 1:
    return data
  import torch
  import torch.distributed as dist
  on_gpu = dist.get_backend() == 'nccl'
  dev = torch.device('cuda', torch.cuda.current_device()) if on_gpu else torch.device('cpu')
  n_local = len(data['action'])
  counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
  dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=dev))
  counts = [int(c.item()) for c in counts]
  starts = np.concatenate([[0], np.cumsum(counts)])
  assert n_total is None or starts[-1] == n_total
  full = {k: (np.empty((starts[-1],) + v.shape[1:], v.dtype) if rank == 0 else None) for k, v in data.items()}
  for k, v in data.items():
    row = int(np.prod(v.shape[1:])) * v.dtype.itemsize
    rows_per_chunk = max(1, chunk_bytes // max(row, 1))
    if rank == 0:
      full[k][:counts[0]] = v
    for src in range(1, world):
      if rank not in (0, src):
        continue
      for lo in range(0, counts[src], rows_per_chunk):
        hi = min(lo + rows_per_chunk, counts[src])
        if rank == src:
          t = torch.from_numpy(np.ascontiguousarray(v[lo:hi]).view(np.uint8).reshape(-1))
          dist.send(t.to(dev), dst=0)
        else:
          t = torch.empty(((hi - lo) * row,), dtype=torch.uint8, device=dev)
          dist.recv(t, src=src)
          full[k][starts[src] + lo:starts[src] + hi] = t.cpu().numpy().view(v.dtype).reshape((hi - lo,) + v.shape[1:])
  dist.barrier()
  return full if rank == 0 else None


def split_seed(seed, split):
  """Stream family of a dataset split.  The reference fills train and test barrels from unseeded RNGs, so they never
  coincide; here rollouts are keyed by (seed, global rollout index), so the split is folded into the key -- otherwise
  `--split=test` would re-create the first train barrels.  'train' keeps the plain seed."""
  if split in (None, '', 'train'):
    return int(seed)
  import zlib
  return (int(seed) ^ (zlib.crc32(str(split).encode()) << 32)) & 0xFFFFFFFFFFFFFFFF


def collect_arrays(env, n_rollouts, T, batch=65536, seed=0, device=None, world_offset=0, progress=None, vec=None):
  """n_rollouts x T random-action rollouts of `env` -> dict of host arrays in the reference layout.  `vec`: an existing
  VecWorldEnv to re-key and reuse (kept open for the caller) instead of allocating one."""
  import torch
  from boxlcd_b200.vec_env import VecWorldEnv
  S, A, P = env.obs_size, env.act_size, max(env.pobs_size, 1)
  H, W = env.observation_space.spaces['lcd'].shape
  # np.empty: every element is written below; the copies land straight in these arrays (no intermediate host tensors)
  out = {'action': np.empty((n_rollouts, T, A), np.float64), 'full_state': np.empty((n_rollouts, T, S), np.float32),
         'proprio': np.empty((n_rollouts, T, P), np.float32), 'lcd': np.empty((n_rollouts, T, H, W), np.bool_)}
  pidx = torch.as_tensor(np.asarray(env.pobs_idxs, np.int64)) if env.pobs_size else None
  done, t0 = 0, time.time()
  own = vec is None
  r = None
  while done < n_rollouts:
    n = min(batch, n_rollouts - done)
    if vec is not None and vec.n == n:   # same batch size: keep the allocation (state, staging, output tensors), move the RNG window
      vec.rekey(seed, world_offset + done)
    else:
      if vec is not None and own:
        vec.close()
      vec, r, own = VecWorldEnv(env, n, device=device, seed=seed, world_offset=world_offset + done), None, True
    vec.reset_dev()
    r = vec.rollout_dev(T) if r is None else vec.rollout_dev(T, r['full_state'], r['lcd_bits'], r['action'])
    # conversions (f32 -> f64 actions, proprio gather, bit unpacking) run on the GPU; each result is copied directly into
    # its slice of the output array
    sl = slice(done, done + n)
    torch.from_numpy(out['action'][sl]).copy_(r['action'].double())
    torch.from_numpy(out['full_state'][sl]).copy_(r['full_state'])
    if pidx is not None:
      torch.from_numpy(out['proprio'][sl]).copy_(r['full_state'].index_select(2, pidx.to(r['full_state'].device)))
    else:
      out['proprio'][sl] = 0.0
    for lo in range(0, n, 8192):   # frames are unpacked in slices: one byte per pixel of transient device memory
      torch.from_numpy(out['lcd'][done + lo:done + min(lo + 8192, n)]).copy_(vec.unpack_lcd(r['lcd_bits'][lo:lo + 8192]))
    done += n
    if progress:
      progress(done, n_rollouts, done * T / (time.time() - t0))
  if vec is not None and own:
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
    # Each rank owns a contiguous range of barrels and simulates them in groups of up to `batch` rollouts per launch
    # (a 1000-world launch would leave most of the GPU idle); barrel b holds global rollouts [1000 b, 1000 (b + 1)) of the
    # split's stream family whatever the grouping or the rank count.  A writer thread deflates group g while group g + 1
    # is being simulated.
    from concurrent.futures import ThreadPoolExecutor
    seed = split_seed(G.seed, G.split)
    b_lo, b_hi = shard_range(G.barrels, rank, world)
    group = max(1, min(int(G.batch) // BARREL_SIZE, 32))
    stages = {'simulate_s': 0.0, 'write_wait_s': 0.0}

    def write_group(first, data):
      for k in range(len(data['action']) // BARREL_SIZE):
        sl = slice(k * BARREL_SIZE, (k + 1) * BARREL_SIZE)
        stamp = datetime.now().strftime('%Y%m%dT%H%M%S') + (f'r{rank}b{first + k}' if world > 1 or G.barrels > 1 else '')
        savez_compressed_parallel(logdir / f'{stamp}-{T}.barrel', **{key: val[sl] for key, val in data.items()})

    pending = None
    with ThreadPoolExecutor(1) as pool:
      for first in range(b_lo, b_hi, group):
        nb = min(group, b_hi - first)
        t0 = time.time()
        data = collect_arrays(env, nb * BARREL_SIZE, T, nb * BARREL_SIZE, seed, world_offset=first * BARREL_SIZE, progress=say)
        stages['simulate_s'] += time.time() - t0
        t0 = time.time()
        if pending is not None:
          pending.result()
        stages['write_wait_s'] += time.time() - t0
        pending = pool.submit(write_group, first, data)
      t0 = time.time()
      if pending is not None:
        pending.result()
      stages['write_wait_s'] += time.time() - t0
    if rank == 0:
      print(f'\nrank 0: {b_hi - b_lo} barrels x {BARREL_SIZE} rollouts x {T} steps: simulate + device->host {stages["simulate_s"]:.2f} s, '
            f'waiting for the barrel writer {stages["write_wait_s"]:.2f} s')
  else:
    N = G.collect_n
    lo, hi = shard_range(N, rank, world)
    data = collect_arrays(env, hi - lo, T, G.batch, G.seed, world_offset=lo, progress=say)
    data = gather_shards(data, rank, world, N)
    if rank == 0:
      os.makedirs('rollouts', exist_ok=True)
      savez_compressed_parallel(f'rollouts/{G.env}-{N}.npz', **data)
      print(f'\nwrote rollouts/{G.env}-{N}.npz')
  if world > 1:
    import torch.distributed as dist
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
