"""Pins the scene constants, the observation / action layout (Appendix C) and the reset algebra and
ranges (Appendix B) against the UNMODIFIED reference run under recording stubs (tests/golden/make_reset_golden.py)."""
import json
import os
import numpy as np
import pytest
import boxlcd_b200 as blcd
from boxlcd_b200 import spec as S
from oracle import oracle

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'reset_golden.json')))
ENVS = sorted(GOLD)
f32 = lambda x: float(np.float32(x))


@pytest.mark.parametrize('name', ENVS)
def test_metadata_matches_reference_init(name):
  g = GOLD[name]
  env = blcd.env_map[name]()
  assert env.obs_keys == g['obs_keys'] and env.act_keys == g['act_keys'] and env.pobs_idxs == g['pobs_idxs']
  assert (env.obs_size, env.act_size, env.pobs_size) == (g['obs_size'], g['act_size'], g['pobs_size'])
  assert env.WIDTH == g['WIDTH'] and env.HEIGHT == g['HEIGHT']
  assert list(env.observation_space.spaces['lcd'].shape) == g['lcd_shape']
  for k, v in g['ENV_DG'].items():
    assert env.G[k] == v, k
  assert env.observation_space.spaces['full_state'].shape == (g['obs_size'],)
  assert env.observation_space.spaces['proprio'].shape == (max(g['pobs_size'], 1),)
  assert env.action_space.shape == (g['act_size'],)


@pytest.mark.parametrize('name', ENVS)
def test_compiled_spec_matches_what_the_reference_hands_to_box2d(name):
  g = GOLD[name]
  sp = blcd.env_map[name]().layout.spec
  first = g['resets'][0]
  assert sp.n_bodies == len(first['bodies']) and sp.n_joints == len(first['joints'])
  for b, gb in enumerate(first['bodies']):
    bd = sp.bodies[b]
    assert f32(bd.density) == f32(gb['density']) and f32(bd.friction) == f32(gb['friction']) and f32(bd.restitution) == f32(gb['restitution']), gb['name']
    assert bd.category_bits == gb['categoryBits'] and bd.mask_bits == gb['maskBits'], gb['name']
    assert f32(bd.linear_damping) == f32(gb['linearDamping']) and f32(bd.angular_damping) == f32(gb['angularDamping'])
    if bd.n_variants == 1:
      sd = bd.shape[0]
      if gb['shape']['kind'] == 'circle':
        assert sd.kind == S.SHAPE_CIRCLE and f32(sd.radius) == f32(gb['shape']['radius'])
      elif sd.kind == S.SHAPE_BOX:
        hx, hy = f32(sd.verts[0][0]), f32(sd.verts[0][1])
        assert [[f32(x), f32(y)] for x, y in gb['shape']['vertices']] == [[-hx, -hy], [hx, -hy], [hx, hy], [-hx, hy]]
      else:
        pts = [(f32(sd.verts[i][0]), f32(sd.verts[i][1])) for i in range(sd.n_verts)]
        order = S.hull_order(pts)
        assert [[f32(x), f32(y)] for x, y in gb['shape']['vertices']] == [list(pts[i]) for i in order]
  for j, gj in enumerate(first['joints']):
    jd = sp.joints[j]
    assert (jd.body_a, jd.body_b) == (gj['bodyA'], gj['bodyB'])
    assert [jd.anchor_a[0], jd.anchor_a[1]] == gj['localAnchorA'] and [jd.anchor_b[0], jd.anchor_b[1]] == gj['localAnchorB']
    assert bool(jd.enable_motor) == gj['enableMotor'] and bool(jd.enable_limit) == gj['enableLimit']
    assert jd.max_motor_torque == gj['maxMotorTorque'] and jd.lower == gj['lowerAngle'] and jd.upper == gj['upperAngle'] and gj['motorSpeed'] == 0


def _chain(sp, b):
  """bodies from b up to (excluding) its robot root"""
  while sp.bodies[b].parent >= 0:
    yield b
    b = sp.bodies[b].parent


@pytest.mark.parametrize('name', ENVS)
def test_child_placement_algebra_matches_reference(name):
  g = GOLD[name]
  sp = blcd.env_map[name]().layout.spec
  for item in g['resets']:
    pose_in = np.array([[b['position'][0], b['position'][1], b['angle64']] for b in item['bodies']], np.float64)
    placed = oracle.place_children(sp, pose_in)
    ref = np.array([[b['position'][0], b['position'][1], b['angle']] for b in item['bodies']], np.float32)
    # the recording stub adds b2Vec2 + ndarray in float64 and rounds once; pybox2d rounds each operand to float32 first
    # (1 ulp at 5 m is 4.8e-7, at 10 m 9.5e-7; a three-link chain accumulates up to two of them, the crab's five-link
    # arm + claw chain up to four)
    depth = max(sum(1 for _ in _chain(sp, b)) for b in range(sp.n_bodies))
    assert np.abs(placed - ref).max() <= (1.5e-6 if depth <= 3 else 1.0e-6 * depth), (name, item["seed"])


@pytest.mark.parametrize('name', ENVS)
def test_reset_ranges_match_reference_samples(name):
  """our resets (own RNG) must fall in the ranges the reference's samples span, body by body"""
  g = GOLD[name]
  env = blcd.env_map[name]()
  ref = np.array([[[b['position'][0], b['position'][1], b['angle']] for b in item['bodies']] for item in g['resets']])
  ow = oracle.OracleWorlds(env.layout.spec, 4096, seed=1, threads=4)
  ow.reset()
  ours = ow.get_bodies()[..., :3]
  for b in range(ref.shape[1]):
    role = env.layout.spec.bodies[b].role
    if role == S.ROLE_CHILD:
      continue
    assert ours[:, b, 0].min() <= ref[:, b, 0].min() + 1e-6 and ours[:, b, 0].max() >= ref[:, b, 0].max() - 1e-6
    assert ours[:, b, 1].min() <= ref[:, b, 1].min() + 1e-6 and ours[:, b, 1].max() >= ref[:, b, 1].max() - 1e-6
    if role == S.ROLE_ROOT:
      assert np.unique(ref[:, b, 1]).size == 1 and (ours[:, b, 1] == np.float32(ref[0, b, 1])).all()   # constant spawn height
    # same support: uniform over [lo, hi] with the reference's bounds
    ext, W, H = env.layout.spec.bodies[b].extent, env.WIDTH, env.HEIGHT
    assert ours[:, b, 0].min() >= ext - 1e-6 and ours[:, b, 0].max() <= W - ext + 1e-6
    assert ref[:, b, 0].min() >= ext - 1e-6 and ref[:, b, 0].max() <= W - ext + 1e-6


def test_reference_observation_of_reset_matches_our_observe_formula():
  # full_state of the reference's reset == rmapto(position), cos/sin(angle) in our layout (Appendix C)
  g = GOLD['LuxoCube']
  env = blcd.env_map['LuxoCube']()
  for item in g['resets'][:4]:
    fs = np.zeros(env.obs_size)
    for b, gb in enumerate(item['bodies']):
      oi = env.layout.spec.bodies[b].obs_index
      fs[oi[0]] = 2 * gb['position'][0] / env.WIDTH - 1
      fs[oi[1]] = 2 * gb['position'][1] / env.HEIGHT - 1
      fs[oi[2]], fs[oi[3]] = np.cos(gb['angle']), np.sin(gb['angle'])
    assert np.abs(fs - np.array(item['full_state'])).max() < 1e-12
