"""GPU tests of the library's large-scene profile (boxlcd_b200/csrc/blcd_profile.h): Crab / CrabCube / SpiderCube
(envs.py:116-137: up to 18 bodies, 64 x 32 frames) and the equivalence of the two profiles on scenes both can run."""
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


@pytest.mark.parametrize('name', ['Urchin', 'LuxoCubes', 'Object3'])
def test_small_scenes_give_identical_results_on_both_profiles(name, monkeypatch):
  env = make_env(name)
  n, T = 1024, 20
  out = {}
  for prof in ('small', 'large'):
    monkeypatch.setenv('BLCD_PROFILE', prof)
    v = vec(env, n, seed=3)
    assert v.info()['profile'] == (1 if prof == 'large' else 0)
    v.reset_dev()
    r = v.rollout_dev(T)
    out[prof] = {k: t.cpu().numpy() for k, t in r.items()}
    out[prof]['counters'] = v.counters()
    v.close()
  for k in out['small']:
    assert (out['small'][k] == out['large'][k]).all(), k


def test_auto_profile_choice():
  assert vec(make_env('UrchinCubes'), 8).info()['profile'] == 0
  v = vec(make_env('Crab'), 8)
  info = v.info()
  assert info['profile'] == 1 and info['n_bodies'] == 17 and info['n_joints'] == 16 and (info['lcd_w'], info['lcd_h']) == (64, 32)


def test_wide_frames_at_several_sizes_match_oracle():
  env = make_env('CrabCube')
  sp = env.layout.spec
  ow = oracle.OracleWorlds(sp, 256, seed=4, threads=8)
  ow.reset()
  ow.rollout(20, want=())
  poses, _ = ow.get_poses()
  v = vec(env, 1)
  for (w, h) in [(64, 32), (48, 24), (40, 16), (32, 16), (64, 64), (33, 8)]:
    ref = oracle.lcd_render(ow.lcd_shapes(0), poses, env.WIDTH, w, h)
    bits = v.render_poses_dev(torch.as_tensor(poses).cuda(), None, w, h).cpu().numpy().view(np.uint32)
    assert bits.shape == ref.shape and (bits == ref).all(), (w, h)


def test_crabcube_numpy_api_and_host_step():
  """the reference-facing call shapes on a 64 x 32 env: dict obs with lcd [N, 32, 64] bool, 12 actions"""
  env = make_env('CrabCube')
  n = 64
  v = vec(env, n, seed=1)
  obs = v.reset()
  assert obs['lcd'].shape == (n, 32, 64) and obs['lcd'].dtype == np.bool_ and obs['full_state'].shape == (n, 72) and obs['proprio'].shape == (n, 68)
  act = np.random.RandomState(0).uniform(-1, 1, (n, 12)).astype(np.float32)
  obs2, rew, done, info = v.step(act)
  assert obs2['lcd'].shape == (n, 32, 64) and not done.any()
  ow = oracle.OracleWorlds(env.layout.spec, n, seed=1, threads=4)
  ow.reset()
  ow.step(act)
  oo = ow.observe()
  assert (np.abs(obs2['full_state'] - oo['full_state']).max(1) < 5e-5).mean() > 0.95
  assert (oracle.unpack_bits(oo['lcd_bits'], 64) == obs2['lcd']).all((1, 2)).mean() > 0.9
  # host-buffer entry point: [N, 32, 2] words per frame
  import ctypes as C
  fs = np.zeros((n, 72), np.float32); bits = np.zeros((n, 32, 2), np.uint32); dn = np.zeros(n, np.uint8)
  p = lambda a: a.ctypes.data_as(C.c_void_p)
  assert v.l.blcd_step_host(v.h, p(act), p(fs), p(bits), p(dn)) == 0
  o3 = v.observe()
  assert (fs == o3['full_state']).all() and (oracle.unpack_bits(bits, 64) == o3['lcd']).all()
  assert v.counters()[:, 5].sum() == 0


def test_crab_collect_dataset_format(tmp_path):
  from boxlcd_b200 import collect
  arrs = collect.collect_arrays(make_env('Crab'), 32, 100, seed=0)
  assert arrs['lcd'].shape == (32, 100, 32, 64) and arrs['lcd'].dtype == np.bool_
  assert arrs['full_state'].shape == (32, 100, 68) and arrs['action'].shape == (32, 100, 12) and arrs['action'].dtype == np.float64
  ink = (~arrs['lcd']).sum((2, 3))
  assert ink.min() > 40 and ink.max() < 700     # a crab is always on screen: shell + 16 limbs at 6.4 px / m
