"""N > 1 host logic on CPU: two gloo ranks each simulate their contiguous shard of worlds (through the host build of the
CUDA path's simulation source -- there is no GPU here), gather on rank 0 with boxlcd_b200.collect.gather_shards, and the
result must equal the single-process run: worlds are keyed by GLOBAL index, so the data cannot depend on the rank count."""
import os
import sys
import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, n_total, T, out_path):
  sys.path.insert(0, HERE)
  sys.path.insert(0, os.path.dirname(HERE))
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  import torch.distributed as dist
  import boxlcd_b200 as blcd
  from boxlcd_b200.collect import shard_range, gather_shards
  from hostsim_py import HostSim
  dist.init_process_group('gloo', rank=rank, world_size=world)
  env = blcd.envs.Urchin()
  lo, hi = shard_range(n_total, rank, world)
  hs = HostSim(env.layout.spec, hi - lo, seed=4, world_offset=lo)
  hs.reset()
  r = hs.rollout(T)
  # small chunks: every array crosses in several point-to-point messages
  data = gather_shards({'full_state': r['full_state'], 'lcd_bits': r['lcd_bits'], 'action': r['action']}, rank, world, n_total, chunk_bytes=700)
  if rank == 0:
    np.savez(out_path, **data)
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_collection_equals_single_process(tmp_path):
  sys.path.insert(0, HERE)
  import boxlcd_b200 as blcd
  from hostsim_py import HostSim, lib
  lib()  # build the host library once, before forking
  n_total, T = 10, 6
  out = str(tmp_path / 'gathered.npz')
  mp.spawn(_worker, args=(2, 29533 + os.getpid() % 500, n_total, T, out), nprocs=2, join=True)
  got = np.load(out)
  env = blcd.envs.Urchin()
  hs = HostSim(env.layout.spec, n_total, seed=4)
  hs.reset()
  ref = hs.rollout(T)
  for k in ('full_state', 'lcd_bits', 'action'):
    assert got[k].shape == ref[k].shape and (got[k] == ref[k]).all(), k


def test_shard_ranges_partition_exactly():
  from boxlcd_b200.collect import shard_range
  for n in (0, 1, 7, 1000, 262144):
    for world in (1, 2, 3, 8):
      spans = [shard_range(n, r, world) for r in range(world)]
      assert spans[0][0] == 0 and spans[-1][1] == n
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
      assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_split_is_part_of_the_stream_key():
  """with one seed, test barrels must not re-create the first train barrels (the reference's are independent draws)"""
  from boxlcd_b200.collect import split_seed
  assert split_seed(0, 'train') == 0 and split_seed(5, None) == 5
  assert split_seed(0, 'test') != split_seed(0, 'train') and split_seed(0, 'test') == split_seed(0, 'test')
  assert split_seed(0, 'test') != split_seed(1, 'test') and 0 <= split_seed(3, 'val') < 2 ** 64
  import boxlcd_b200 as blcd
  from hostsim_py import HostSim
  env = blcd.envs.Urchin()
  a = HostSim(env.layout.spec, 4, seed=split_seed(0, 'train')); a.reset()
  b = HostSim(env.layout.spec, 4, seed=split_seed(0, 'test')); b.reset()
  assert not (a.rollout(3)['action'] == b.rollout(3)['action']).any()
