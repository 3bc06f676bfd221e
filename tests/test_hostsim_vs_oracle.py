"""The CUDA path's per-world simulation source (boxlcd_b200/csrc/blcd_world.cuh) compiled for the host must reproduce
the CPU oracle BIT FOR BIT: same fp32 operation order, same contact ordering, same RNG.  This is the logic check that
can run without a GPU; the GPU tests then only have to absorb sincosf / FMA-level differences."""
import numpy as np
import pytest
import boxlcd_b200 as blcd
from oracle import oracle
from hostsim_py import HostSim
from common import make_env, random_bodies

ALL = sorted(blcd.env_map)


@pytest.mark.parametrize('name', ALL)
def test_rollouts_bit_exact(name):
  env = make_env(name)
  n, T = 24, 40
  ow = oracle.OracleWorlds(env.layout.spec, n, seed=11, threads=4)
  hs = HostSim(env.layout.spec, n, seed=11)
  ow.reset(); hs.reset()
  assert (ow.get_bodies() == hs.get_bodies()).all()
  ro, rh = ow.rollout(T), hs.rollout(T)
  for k in ro:
    assert (ro[k] == rh[k]).all(), k
  co, ch = ow.counters(), hs.counters()
  assert ch[:, oracle.COUNTER_NAMES.index('overflow')].sum() == 0
  assert (co == ch).all()


@pytest.mark.parametrize('name', ['Urchin', 'LuxoCube', 'UrchinBall', 'Object2', 'Bounce2', 'CrabCube', 'SpiderCube'])
def test_single_steps_from_fresh_states_bit_exact(name):
  env = make_env(name)
  rng = np.random.RandomState(5)
  n = 96
  bodies, variants = random_bodies(env, n, rng)
  act = rng.uniform(-1.5, 1.5, (n, env.act_size)).astype(np.float32)
  ow = oracle.OracleWorlds(env.layout.spec, n)
  hs = HostSim(env.layout.spec, n)
  ow.set_bodies(bodies, variants); hs.set_bodies(bodies, variants)
  for _ in range(2):
    ow.step(act); hs.step(act)
    assert (ow.get_bodies() == hs.get_bodies()).all()
  oo, oh = ow.observe(), hs.observe()
  assert (oo['full_state'] == oh['full_state']).all() and (oo['lcd_bits'] == oh['lcd_bits']).all()


def test_reset_from_full_state_bit_exact():
  env = make_env('LuxoCube')
  n = 32
  ow = oracle.OracleWorlds(env.layout.spec, n, seed=1); ow.reset()
  fs = ow.observe()['full_state']
  ow2 = oracle.OracleWorlds(env.layout.spec, n, seed=2); hs2 = HostSim(env.layout.spec, n, seed=2)
  ow2.reset(full_state=fs); hs2.reset(full_state=fs)
  assert (ow2.get_bodies() == hs2.get_bodies()).all()
  ow2.step(np.zeros((n, 3), np.float32)); hs2.step(np.zeros((n, 3), np.float32))
  assert (ow2.get_bodies() == hs2.get_bodies()).all()


def test_host_rasterizer_matches_reference_golden_frames():
  import os
  gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'lcd_golden.npz'))
  for name in ['Dropbox', 'Bounce2', 'Object2', 'Urchin', 'Luxo', 'UrchinCube', 'LuxoCube', 'UrchinBall', 'LuxoBall', 'Crab', 'SpiderCube']:
    env = make_env(name)
    hs = HostSim(env.layout.spec, 1)
    kind = gold[f'{name}_kind']
    variants = (kind != 0).astype(np.uint32) @ (1 << np.arange(kind.shape[1])).astype(np.uint32)
    if not any(env.layout.spec.bodies[b].n_variants > 1 for b in range(kind.shape[1])):
      variants = None
    bits = hs.render_poses(gold[f'{name}_poses'], variants)
    assert bits.shape == gold[f'{name}_bits'].shape and (bits == gold[f'{name}_bits']).all(), name


@pytest.mark.parametrize('profile', ['small', 'large'])
def test_host_rasterizer_at_the_human_view_size(profile):
  """frames wider than one row mask are rendered as column windows (32 px small profile, 64 px large): x8 reference frames"""
  import os
  gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'lcd_golden.npz'))
  for name in ['Urchin', 'LuxoCube']:
    hs = HostSim(make_env(name).layout.spec, 1, profile=profile)
    _, w8, h8 = [int(x) for x in gold[f'{name}_x8_meta']]
    bits = hs.render_poses(gold[f'{name}_x8_poses'], None, w8, h8)
    assert bits.shape == gold[f'{name}_x8_bits'].shape and (bits == gold[f'{name}_x8_bits']).all(), name


@pytest.mark.parametrize('name', ['Urchin', 'LuxoCubes', 'Object3', 'UrchinBall'])
def test_large_profile_build_is_bit_identical_on_small_scenes(name):
  """the two scene-size profiles (csrc/blcd_profile.h) are the same source with different limits: same results"""
  env = make_env(name)
  n, T = 16, 40
  a, b = HostSim(env.layout.spec, n, seed=4, profile='small'), HostSim(env.layout.spec, n, seed=4, profile='large')
  a.reset(); b.reset()
  ra, rb = a.rollout(T), b.rollout(T)
  for k in ra:
    assert (ra[k] == rb[k]).all(), k
  assert (a.counters() == b.counters()).all()


@pytest.mark.parametrize('name', ['Urchin', 'LuxoCube', 'SpiderCube'])
def test_open_world_floor_only_bit_exact(name):
  """walls=0 (world_env.py:316): one long floor edge, bodies may leave the frame"""
  env = make_env(name, walls=0)
  sp = env.layout.spec
  assert sp.n_walls == 1 and tuple(sp.walls[0]) == (-1000.0 * env.WIDTH, 0.0, 1000.0 * env.WIDTH, 0.0)
  n, T = 24, 60
  ow, hs = oracle.OracleWorlds(sp, n, seed=2, threads=4), HostSim(sp, n, seed=2)
  ow.reset(); hs.reset()
  ro, rh = ow.rollout(T), hs.rollout(T)
  for k in ro:
    assert (ro[k] == rh[k]).all(), k
  b = ow.get_bodies()
  assert np.isfinite(b).all() and b[..., 1].min() > 0.0          # nothing falls through the floor
  assert (b[..., 0] < 0).any() or (b[..., 0] > env.WIDTH).any()    # and nothing keeps the robots inside [0, WIDTH]


@pytest.mark.parametrize('name', ALL)
def test_phase_pipeline_rollouts_bit_exact(name):
  """csrc/blcd_pipeline.cuh: the sub-step as five phases (setup / velocity / position / finish / TOI) with the records passed
  through a per-world scratch buffer and a fresh, poisoned simulator object per phase -- the data flow between the
  pipeline's kernels -- must reproduce the oracle bit for bit, counters included"""
  env = make_env(name)
  n, T = 24, 40
  ow = oracle.OracleWorlds(env.layout.spec, n, seed=13, threads=4)
  hs = HostSim(env.layout.spec, n, seed=13)
  ow.reset(); hs.reset()
  ro, rh = ow.rollout(T), hs.rollout(T, pipeline=True)
  for k in ro:
    assert (ro[k] == rh[k]).all(), k
  assert (ow.get_bodies() == hs.get_bodies()).all()
  assert (ow.counters() == hs.counters()).all()


@pytest.mark.parametrize('name', ['Urchin', 'LuxoCube', 'Object2', 'CrabCube'])
def test_phase_pipeline_single_steps_and_mixing_with_the_fused_path(name):
  env = make_env(name)
  rng = np.random.RandomState(6)
  n = 64
  bodies, variants = random_bodies(env, n, rng)
  act = rng.uniform(-1.5, 1.5, (n, env.act_size)).astype(np.float32)
  ow, hs = oracle.OracleWorlds(env.layout.spec, n), HostSim(env.layout.spec, n)
  ow.set_bodies(bodies, variants); hs.set_bodies(bodies, variants)
  for t in range(4):     # alternate the two paths on the same worlds: the persistent state is all that carries over
    ow.step(act); hs.step(act, pipeline=(t % 2 == 0))
    assert (ow.get_bodies() == hs.get_bodies()).all(), t
  assert (ow.counters() == hs.counters()).all()


@pytest.mark.parametrize('profile', ['small', 'large'])
def test_converged_warp_scanline_rules_equal_the_row_rules(profile):
  """csrc/blcd_raster.cuh: polygon_row_t (unrolled edges, sorting network, per-body slope table, integer rounding forms) --
  what the render kernel runs -- against polygon_row (pinned to the reference through the golden frames) on 3 million random
  integer polygons incl. degenerate ones, both Pillow rule sets, column windows, rows outside the polygon"""
  from hostsim_py import lib
  assert lib(large=(profile == 'large')).hostsim_polygon_row_check(3_000_000, 7) == 0


@pytest.mark.parametrize('robot', ['quad', 'legs', 'walker', 'gingy', 'octo'])
def test_robot_fillers_without_a_named_env_build_and_simulate(robot):
  """world_defs.py:125-167, 251-368: robots the reference defines but never wires into a named env.  A custom WorldDef with each
  must compile to a scene, reset, and step bit-for-bit like the oracle (small or large profile, whichever it needs), with
  every hinge holding and nothing leaving the arena."""
  from boxlcd_b200.world_env import WorldEnv
  from boxlcd_b200.world_defs import WorldDef, Robot, Object
  objects = [Object('object0', shape='box', size=0.4, density=0.5)] if robot in ('quad', 'walker') else []
  env = WorldEnv(WorldDef(robots=[Robot(type=robot, name=f'{robot}0')], objects=objects), {'lcd_base': 32 if robot in ('gingy', 'octo', 'walker') else 16})
  sp = env.layout.spec
  assert sp.n_joints >= 2 and sp.n_bodies == sp.n_joints + 1 + len(objects)
  assert env.act_size == sum(1 for j in range(sp.n_joints) if sp.joints[j].act_index >= 0) and env.obs_size == 4 * sp.n_bodies
  n, T = 8, 25
  ow, hs = oracle.OracleWorlds(sp, n, seed=3, threads=4), HostSim(sp, n, seed=3)
  ow.reset(); hs.reset()
  assert (ow.get_bodies() == hs.get_bodies()).all()
  ro, rh = ow.rollout(T), hs.rollout(T)
  for k in ro:
    assert (ro[k] == rh[k]).all(), k
  rp = HostSim(sp, n, seed=3); rp.reset()
  assert (rp.rollout(T, pipeline=True)['full_state'] == ro['full_state']).all()
  b = ow.get_bodies()
  assert np.isfinite(b).all() and b[..., 1].min() > -0.2 and b[..., 0].min() > -0.2 and b[..., 0].max() < env.WIDTH + 0.2
  assert (~oracle.unpack_bits(ro['lcd_bits'], sp.lcd_w)).sum((2, 3)).min() >= 8      # the robot is on the frame
