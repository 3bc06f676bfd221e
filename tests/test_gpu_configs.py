"""BASELINE.json's configs as GPU test cases (full sizes where the check is a size-independent property), the dataset
formats of examples/collect.py / research/data.py, and edge cases (ragged batch sizes, empty / partial resets, errors)."""
import os
import numpy as np
import pytest
import torch
import boxlcd_b200 as blcd
from oracle import oracle
from common import make_env

pytestmark = pytest.mark.gpu


def vec(env, n, **kw):
  from boxlcd_b200.vec_env import VecWorldEnv
  return VecWorldEnv(env, n, **kw)


def check_world_invariants(env, v, slack=0.12):
  b = v.get_bodies()
  assert np.isfinite(b).all()
  assert (b[..., 0] > -slack).all() and (b[..., 0] < env.WIDTH + slack).all()
  assert (b[..., 1] > -slack).all() and (b[..., 1] < env.HEIGHT + slack).all()
  assert v.counters()[:, 5].sum() == 0, 'manifold slots overflowed'
  return b


def frames_match_oracle_rasterizer(env, v, sample=4096, seed=0):
  """frames written by the fused kernel == oracle rasterizer applied to the kernel's own transforms (bit-exact)"""
  obs = v.observe_dev()
  bits = obs['lcd_bits'].cpu().numpy().view(np.uint32)
  poses, variants = v.get_poses_dev()
  idx = np.random.RandomState(seed).choice(v.n, min(sample, v.n), replace=False)
  ow = oracle.OracleWorlds(env.layout.spec, 1)
  ow.reset()
  shapes = ow.lcd_shapes(0)
  if any(env.layout.spec.bodies[k].n_variants > 1 for k in range(v.B)):
    pytest.skip('per-world shapes')
  ref = oracle.lcd_render(shapes, poses.cpu().numpy()[idx], env.WIDTH, v.W, v.H)
  assert (bits[idx] == ref).all()


def test_config0_dropbox_200_step_collect_format(tmp_path, monkeypatch):
  """envs.Dropbox() 16x16, random actions, 200-step rollouts via collect (examples/collect.py:24-41)"""
  from boxlcd_b200 import collect
  monkeypatch.chdir(tmp_path)
  collect.main(['--env=Dropbox', '--collect_n=300', '--ep_len=200', '--batch=128'])
  d = np.load(tmp_path / 'rollouts' / 'Dropbox-300.npz')
  assert set(d.files) == {'action', 'full_state', 'proprio', 'lcd'}
  assert d['action'].shape == (300, 200, 1) and d['action'].dtype == np.float64
  assert d['full_state'].shape == (300, 200, 4) and d['full_state'].dtype == np.float32
  assert d['proprio'].shape == (300, 200, 1) and d['proprio'].dtype == np.float32 and (d['proprio'] == 0).all()
  assert d['lcd'].shape == (300, 200, 16, 16) and d['lcd'].dtype == np.bool_
  assert d['action'].min() >= -1 and d['action'].max() <= 1
  ink = (~d['lcd']).sum((2, 3))
  assert ink.min() >= 16 and ink.max() <= 40        # the 1.4 m (4.48 px) box covers 18..36 px depending on its angle
  last = ~d['lcd'][:, -1]
  rows = last.any(2)
  assert (rows[:, :9] == False).all() and rows[:, 11:].all()   # at rest on the floor: rows 11-15
  assert np.isin(last.sum((1, 2)), [24, 25, 30]).mean() > 0.9  # flat on the floor: a 5x5 or 5x6 block (oracle: 99 %)
  # element [i, j] is the observation BEFORE action j: consecutive observations differ while the box is falling
  assert (np.abs(np.diff(d['full_state'][:, :5], axis=1)).max(2) > 0).any()


def test_config1_bounce2_65k_worlds():
  env = make_env('Bounce2')
  v = vec(env, 65536, seed=2)
  v.reset_dev()
  r = v.rollout_dev(50)
  b = check_world_invariants(env, v)
  bits = r['lcd_bits'].cpu().numpy().view(np.uint32)
  ink = (~oracle.unpack_bits(bits[::64], 16)).sum((2, 3))
  assert ink.min() >= 8 and ink.max() <= 44 and 20 < ink.mean() < 30   # two r=0.5 balls, 12-16 px each (gif: 24.5 mean)
  frames_match_oracle_rasterizer(env, v)
  # balls do not overlap by more than the solver's slop
  d = np.hypot(b[:, 0, 0] - b[:, 1, 0], b[:, 0, 1] - b[:, 1, 1])
  assert (d > 1.0 - 0.05).mean() > 0.999


def test_config2_urchin_262k_worlds():
  env = make_env('Urchin')
  v = vec(env, 262144, seed=3)
  v.reset_dev()
  v.rollout_dev(10)
  b = check_world_invariants(env, v)
  for leg in (1, 2, 3):   # revolute anchors stay pinned to the root
    ax = b[:, leg, 0] - np.sin(b[:, leg, 2]) * (20 / 30)
    ay = b[:, leg, 1] + np.cos(b[:, leg, 2]) * (20 / 30)
    err = np.hypot(ax - b[:, 0, 0], ay - b[:, 0, 1])
    # Box2D's TOI sub-steps ignore joints, so a leg that slams a wall can be pulled a few cm off its hinge for a step;
    # the CPU oracle gives the same distribution (97.3-97.9 % below 3 cm, 99.7 % below 10 cm after 10 random steps)
    assert 0.96 < (err < 0.03).mean() < 0.99 and (err < 0.1).mean() > 0.995
  frames_match_oracle_rasterizer(env, v)
  c = v.counters().astype(np.int64)
  assert (c[:, 7] == 30).all()   # 10 env steps x 3 sub-steps for every world


def test_config3_luxocube_barrel_dataset(tmp_path):
  """research/data.py barrel format: 1000 rollouts per file, {timestamp}-{ep_len}.barrel.npz under logdir/<split>/"""
  from boxlcd_b200 import collect
  collect.main(['--env=LuxoCube', '--barrels=2', f'--logdir={tmp_path}', '--split=test'])
  files = sorted((tmp_path / 'test').glob('*.barrel.npz'))
  assert len(files) == 2 and all(f.name.endswith('-150.barrel.npz') for f in files)
  d = np.load(files[0])
  assert d['action'].shape == (1000, 150, 3) and d['action'].dtype == np.float64
  assert d['full_state'].shape == (1000, 150, 20) and d['proprio'].shape == (1000, 150, 16) and d['lcd'].shape == (1000, 150, 16, 24)
  env = make_env('LuxoCube')
  assert (d['proprio'] == d['full_state'][..., env.pobs_idxs]).all()
  d2 = np.load(files[1])
  assert not (d['action'][:10] == d2['action'][:10]).all(), 'barrels must hold different rollouts'
  ink = (~d['lcd']).sum((2, 3))
  assert 38 < ink.mean() < 50   # reference gif: 43.4 mean


@pytest.mark.parametrize('n', [4096, 16384])
def test_config4_urchinball_sweep_sizes(n):
  env = make_env('UrchinBall')
  v = vec(env, n, seed=6)
  v.reset_dev()
  v.rollout_dev(15)
  check_world_invariants(env, v)
  frames_match_oracle_rasterizer(env, v, sample=2048)


@pytest.mark.parametrize('n', [1, 33, 257, 1000])
def test_ragged_batch_sizes_match_oracle(n):
  """batch sizes that leave partial warps / partial blocks (the phase barriers count live warps only)"""
  env = make_env('Urchin')
  v = vec(env, n, seed=8)
  ow = oracle.OracleWorlds(env.layout.spec, n, seed=8, threads=4)
  v.reset_dev(); ow.reset()
  rg = v.rollout_dev(3)
  ro = ow.rollout(3)
  assert (rg['action'].cpu().numpy() == ro['action']).all()
  assert (np.abs(rg['full_state'].cpu().numpy()[:, 1] - ro['full_state'][:, 1]).max(1) < 1e-4).mean() >= 0.95
  assert np.isfinite(v.get_bodies()).all()


def test_partial_and_empty_reset():
  env = make_env('Bounce2')
  v = vec(env, 64, seed=1)
  v.reset_dev()
  v.rollout_dev(5)
  before = v.get_bodies()
  v.reset_dev(torch.zeros(0, dtype=torch.int64, device='cuda'))        # empty index list: nothing happens
  assert (v.get_bodies() == before).all()
  idx = torch.tensor([3, 17, 40], dtype=torch.int64, device='cuda')
  v.reset_dev(idx)
  after = v.get_bodies()
  keep = np.setdiff1d(np.arange(64), [3, 17, 40])
  assert (after[keep] == before[keep]).all() and (after[[3, 17, 40]] != before[[3, 17, 40]]).any()
  assert (after[[3, 17, 40], :, 3:] == 0).all()   # fresh worlds start at rest


def test_errors_are_loud():
  env = make_env('Urchin')
  v = vec(env, 8)
  with pytest.raises(RuntimeError, match='width out of range'):
    v.render(2048, 32)
  with pytest.raises(IndexError):
    blcd.envs.Dropbox({'walls': 0})     # the reference indexes robots[0] for the scroll offset (world_env.py:382)
  with pytest.raises(AssertionError, match='lcd_mode'):
    blcd.envs.Urchin().lcd_render(lcd_mode='L')      # world_env.py:465: only '1' and 'RGB'


def test_object2_random_shapes_render_and_step():
  env = make_env('Object2')
  n = 4096
  v = vec(env, n, seed=5)
  ow = oracle.OracleWorlds(env.layout.spec, n, seed=5, threads=8)
  v.reset_dev(); ow.reset()
  poses, variants = v.get_poses_dev()
  _, var_o = ow.get_poses()
  assert (variants.cpu().numpy().view(np.uint32) == var_o).all()
  assert 0.2 < (var_o & 1).mean() < 0.8        # both shapes occur
  og, oo = v.observe(), ow.observe()
  assert (og['lcd'] == oracle.unpack_bits(oo['lcd_bits'], 16)).all((1, 2)).mean() > 0.999
  v.rollout_dev(20); ow.rollout(20, want=())
  assert np.isfinite(v.get_bodies()).all()


def test_failure_detection_flags_and_resets_non_finite_worlds():
  env = make_env('Bounce2')
  v = vec(env, 64, seed=2)
  v.reset_dev()
  assert v.check_finite()[0] == 0
  b = v.get_bodies()
  b[5, 0, 0] = np.nan
  b[9, 1, 4] = np.inf
  v.set_bodies(b)
  n_bad, flags = v.check_finite()
  assert n_bad == 2 and flags.cpu().numpy().nonzero()[0].tolist() == [5, 9]
  v.check_finite(auto_reset=True)
  assert v.check_finite()[0] == 0 and np.isfinite(v.get_bodies()).all()


def test_open_world_without_walls_matches_oracle_and_tracks_scroll():
  """walls=0 (world_env.py:316, 381-382, 453-454)"""
  env = make_env('Urchin', walls=0)
  n = 2048
  ow = oracle.OracleWorlds(env.layout.spec, n, seed=6, threads=8)
  ow.reset()
  v = vec(env, n, seed=6)
  v.reset_dev()
  assert np.abs(v.get_bodies() - ow.get_bodies()).max() < 2e-6
  act = np.random.RandomState(1).uniform(-1, 1, (n, 3)).astype(np.float32)
  ow.step(act)
  v.step_dev(torch.as_tensor(act).cuda(), observe=False)
  assert (np.abs(v.get_bodies()[..., :3] - ow.get_bodies()[..., :3]).max((1, 2)) < 1e-5).mean() > 0.98
  r = v.rollout_dev(100)
  fs = r['full_state'].cpu().numpy()
  assert np.isfinite(fs).all() and (np.abs(fs[:, -1, env.obs_keys.index('urchin0:root:x:p')]) > 1.0).any()   # left the frame
  e1 = make_env('Urchin', walls=0)
  obs = e1.reset()
  x = (obs['full_state'][e1.obs_keys.index('urchin0:root:x:p')] + 1) / 2 * e1.WIDTH
  assert abs(e1.scroll - (x - e1.WIDTH / 2)) < 1e-6


@pytest.mark.parametrize('n', [300, 5000, 20000])
def test_host_buffer_step_pipeline_matches_device_path(n, monkeypatch):
  """blcd_step_host (host actions in, host obs out; pipelined over world sub-ranges) == step_dev + observe_dev"""
  import ctypes as C
  env = make_env('UrchinBall')
  act = np.random.RandomState(n).uniform(-1, 1, (n, env.act_size)).astype(np.float32)
  p = lambda a: a.ctypes.data_as(C.c_void_p)
  outs = []
  for chunks in ('1', '4', '8'):
    monkeypatch.setenv('BLCD_HOST_CHUNKS', chunks)
    v = vec(env, n, seed=2)
    v.reset_dev()
    fs = np.zeros((n, v.S), np.float32); bits = np.zeros((n, v.H), np.uint32); dn = np.ones(n, np.uint8)
    for _ in range(3):
      assert v.l.blcd_step_host(v.h, p(act), p(fs), p(bits), p(dn)) == 0
    outs.append((fs, bits, dn))
    v.close()
  v = vec(env, n, seed=2)
  v.reset_dev()
  for _ in range(3):
    obs, _ = v.step_dev(torch.as_tensor(act).cuda())
  ref = (obs['full_state'].cpu().numpy(), obs['lcd_bits'].cpu().numpy().view(np.uint32), obs['done'].cpu().numpy())
  for o in outs:
    for a, b in zip(o, ref):
      assert (a == b).all()


def test_async_host_steps_match_synchronous_steps():
  """blcd_step_host_async / _wait (AsyncVectorEnv.step_async / step_wait call shape): a [T, N, ...] dataset filled with up to
  three steps in flight equals the same steps taken one synchronous call at a time"""
  env = make_env('LuxoCube')
  n, T = 6000, 12
  acts = np.random.RandomState(0).uniform(-1, 1, (T, n, env.act_size)).astype(np.float32)
  v = vec(env, n, seed=5)
  fs = np.zeros((T, n, v.S), np.float32); bits = np.zeros((T, n, v.H), np.uint32); dn = np.zeros((T, n), np.uint8)
  with pytest.raises(RuntimeError, match='page-locked'):
    v.step_host_async(acts[0], fs[0], bits[0], dn[0])
  v.pin_host(acts, fs, bits, dn)
  v.reset_dev()
  for t in range(T):
    v.step_host_async(acts[t], fs[t], bits[t], dn[t])
    v.step_host_wait(keep_in_flight=2)
  v.step_host_wait()
  v2 = vec(env, n, seed=5)
  v2.reset_dev()
  fs2 = np.zeros((n, v.S), np.float32); bits2 = np.zeros((n, v.H), np.uint32); dn2 = np.zeros(n, np.uint8)
  for t in range(T):
    v2.step_host(acts[t], fs2, bits2, dn2)
    assert (fs[t] == fs2).all() and (bits[t] == bits2).all() and (dn[t] == dn2).all(), t
  # device-side calls issued afterwards see the state after all T steps
  assert (v.observe()['full_state'] == fs[-1]).all()


def test_vector_env_seed_rekeys_the_world_streams():
  env = make_env('Urchin')
  v = vec(env, 64, seed=1)
  a = v.reset()['full_state']
  v.seed(2)
  b = v.reset()['full_state']
  v.seed([1] * 64)
  c = v.reset()['full_state']
  assert (a == c).all() and (a != b).any()


@pytest.mark.parametrize('name', ['Bounce2', 'Object3', 'Urchin'])
def test_results_do_not_depend_on_the_block_size(name, monkeypatch):
  """blcd_create picks the block size from the scene and the world count; a world's arithmetic must not notice"""
  env = make_env(name)
  n, T = 3000, 25
  outs = []
  for block in ('128', '160', '224', '256', '320', '384', '448', '512'):
    monkeypatch.setenv('BLCD_BLOCK', block)
    v = vec(env, n, seed=9)
    assert v.info()['block'] == int(block)
    v.reset_dev()
    r = v.rollout_dev(T)
    obs, _ = v.step_dev(None)
    outs.append([r[k].cpu().numpy() for k in ('full_state', 'lcd_bits', 'action')] + [obs['full_state'].cpu().numpy(), v.counters()])
    v.close()
  for o in outs[1:]:
    for a, b in zip(outs[0], o):
      assert (a == b).all()
  monkeypatch.delenv('BLCD_BLOCK')
  monkeypatch.delenv('BLCD_PIPELINE', raising=False)    # the automatic path selection is what is checked below
  # wave-aware choices: 65 536 worlds are one wave of 448-thread blocks (also for articulated scenes, since round 2); 32 768
  # worlds (262 144 over 8 GPUs) one wave of 224-thread blocks; large batches go to the phase pipeline
  assert vec(make_env('Bounce2'), 65536).info()['block'] == 448 and vec(make_env('Urchin'), 65536).info()['block'] == 448
  assert vec(make_env('Urchin'), 32768).info()['block'] == 224 and vec(make_env('Urchin'), 32768).info()['pipeline'] == 0
  assert vec(make_env('Urchin'), 131072).info()['pipeline'] == 1
  assert vec(make_env('CrabCube'), 65536).info()['pipeline'] == 1 and vec(make_env('SpiderCube'), 65536).info()['pipeline'] == 1
  assert vec(make_env('SpiderCube'), 32768).info()['pipeline'] == 0


def test_rekeyed_handle_equals_a_fresh_one_and_splits_differ(tmp_path):
  """blcd_rekey (collector hygiene): one allocation walked over a dataset gives exactly what fresh handles would; the split
  is part of the stream key, so test barrels never duplicate train barrels of the same --seed"""
  from boxlcd_b200 import collect
  env = make_env('UrchinBall', ep_len=12)
  a = collect.collect_arrays(env, 700, 12, batch=256, seed=3, world_offset=1000)           # 256 + 256 + 188: rekey, then a new size
  fresh = vec(env, 700, seed=3, world_offset=1000)
  fresh.reset_dev()
  r = fresh.rollout_dev(12)
  assert (a['full_state'] == r['full_state'].cpu().numpy()).all() and (a['action'] == r['action'].cpu().numpy().astype(np.float64)).all()
  assert (a['lcd'] == fresh.unpack_lcd(r['lcd_bits']).cpu().numpy()).all()
  v = vec(env, 64, seed=1)
  v.reset_dev(); x = v.rollout_dev(3)['full_state'].clone()
  v.rekey(1, 0)
  v.reset_dev(); y = v.rollout_dev(3)['full_state'].clone()
  assert torch.equal(x, y), 'rekey to the same key must replay the same streams'
  v.rekey(1, 64)
  v.reset_dev(); z = v.rollout_dev(3)['full_state']
  assert not torch.equal(x, z)
  for split in ('train', 'test'):
    collect.main(['--env=Bounce', '--barrels=1', f'--logdir={tmp_path}', f'--split={split}', '--ep_len=5'])
  tr = np.load(next((tmp_path / 'train').glob('*.barrel.npz')))
  te = np.load(next((tmp_path / 'test').glob('*.barrel.npz')))
  assert tr['full_state'].shape == te['full_state'].shape == (1000, 5, 4)
  assert not (tr['full_state'][:, 0] == te['full_state'][:, 0]).all(1).any(), 'test rollouts must not repeat train rollouts'


def test_step_host_with_fresh_pageable_arrays_every_step():
  """blcd_step_host must be safe with throw-away numpy arrays (no page-locking behind the caller's back): big,
  mmap-backed arrays are allocated and dropped every step, results must equal the device-tensor path"""
  env = make_env('Urchin')
  n = 40000     # [n, 16] f32 = 2.5 MB per array: mmap-backed, unmapped on free
  a, b = vec(env, n, seed=11), vec(env, n, seed=11)
  a.reset_dev(); b.reset_dev()
  rng = np.random.RandomState(0)
  for t in range(6):
    act = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    fs, bits, done = np.empty((n, 16), np.float32), np.empty((n, 16), np.uint32), np.empty(n, np.uint8)
    a.step_host(act.copy(), fs, bits, done)
    obs, _ = b.step_dev(torch.as_tensor(act).cuda(), observe=True)
    assert (fs == obs['full_state'].cpu().numpy()).all() and (bits == obs['lcd_bits'].cpu().numpy().view(np.uint32)).all()
    del fs, bits, done
  # explicit registration still gives the direct-DMA path, and can be undone
  fs = np.empty((n, 16), np.float32)
  a.pin_host(fs)
  a.step_host(act, fs, None, None)
  from boxlcd_b200 import _lib
  _lib.check(a.l.blcd_unpin_host(a.h, fs.ctypes.data))
  with pytest.raises(RuntimeError, match='not registered'):
    _lib.check(a.l.blcd_unpin_host(a.h, fs.ctypes.data))


def test_rgb_mode_and_human_frame_through_the_env_api():
  """lcd_render(lcd_mode='RGB') / render(mode='human', return_pyglet_view=True) (world_env.py:460-535)"""
  env = blcd.envs.UrchinBall()
  env.seed(2)
  env.reset()
  for _ in range(3):
    env.step(env.action_space.sample())
  rgb = env.lcd_render(lcd_mode='RGB')
  assert rgb.shape == (16, 24, 3) and rgb.dtype == np.uint8
  lcd = env.lcd_render()
  # wherever the LCD frame has ink the colour view has a body colour (fill or outline)
  assert ((rgb != 254).any(-1) | lcd).all()
  assert set(map(tuple, rgb.reshape(-1, 3))) <= {(254, 254, 254), (230, 102, 102), (128, 77, 128), (128, 102, 230), (77, 77, 128)}
  img = env.render(mode='human', return_pyglet_view=True)
  assert img.shape == (128, 192 + 1 + 192, 3) and (img[:, 192] == 0).all()
  assert (img[::8, 193::8, 0] == 255 * lcd).all()
  big = env.lcd_render(192, 128, lcd_mode='RGB')
  assert (img[:, :192] == big).all()
  env.close()


@pytest.mark.parametrize('name', ['Urchin', 'LuxoCube', 'Bounce2', 'CrabCube'])
def test_phase_pipeline_and_fused_kernel_agree(name, monkeypatch):
  """The two device paths (csrc/blcd_pipeline.cuh: five kernels per sub-step through HBM scratch; the fused one-thread-per-world
  kernel) run the same phase functions in the same order: identical action streams and counters' structure, states equal up to
  FMA-contraction-level round-off -- each as close to the oracle as the other (tests/test_gpu_parity.py bars)."""
  env = make_env(name)
  n, T = 4096, 4
  outs = {}
  for pipeline in ('0', '1'):
    monkeypatch.setenv('BLCD_PIPELINE', pipeline)
    v = vec(env, n, seed=21)
    v.reset_dev()
    r = v.rollout_dev(T)
    outs[pipeline] = {k: x.cpu().numpy() for k, x in r.items()}
    outs[pipeline]['counters'] = v.counters()
    assert v.counters()[:, 5].sum() == 0
  a, b = outs['0'], outs['1']
  assert (a['action'] == b['action']).all()
  assert (a['full_state'][:, 0] == b['full_state'][:, 0]).all(), 'the reset state does not depend on the path'
  big = env.layout.spec.n_bodies > 8
  close = (np.abs(a['full_state'][:, 1] - b['full_state'][:, 1]).max(1) < (5e-5 if big else 1e-5)).mean()
  assert close > (0.97 if big else 0.99), close
  assert (a['counters'][:, 7] == b['counters'][:, 7]).all()     # sub-steps taken
  ow = oracle.OracleWorlds(env.layout.spec, n, seed=21, threads=8)
  ow.reset()
  ref = ow.rollout(T)['full_state']
  for k in ('0', '1'):
    ok = (np.abs(outs[k]['full_state'][:, 1] - ref[:, 1]).max(1) < (5e-5 if big else 1e-5)).mean()
    assert ok > (0.97 if big else 0.98), (k, ok)


def test_phase_pipeline_is_sharding_invariant(monkeypatch):
  """worlds keyed by global index: a pipeline shard reproduces its slice of the full batch bit for bit, whatever the world ranges /
  streams the launches were split into"""
  monkeypatch.setenv('BLCD_PIPELINE', '1')
  env = make_env('UrchinBall')
  full = vec(env, 3000, seed=9); full.reset_dev(); f = full.rollout_dev(6)
  part = vec(env, 1000, seed=9, world_offset=1500); part.reset_dev(); p = part.rollout_dev(6)
  for k in f:
    assert torch.equal(f[k][1500:2500], p[k]), k
  a = torch.rand((3000, 3), device='cuda') * 2 - 1
  of, _ = full.step_dev(a, observe=True)
  op, _ = part.step_dev(a[1500:2500].contiguous(), observe=True)
  assert torch.equal(of['full_state'][1500:2500], op['full_state']) and torch.equal(of['lcd_bits'][1500:2500], op['lcd_bits'])
