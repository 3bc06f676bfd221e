"""Known-answer and invariant tests that pin the physics oracle (oracle/b2_*.h) where pybox2d itself cannot be run:
analytic answers of Box2D's integrator / solver on configurations simple enough to work out by hand."""
import math
import numpy as np
import pytest
import boxlcd_b200 as blcd
from oracle import oracle

H = np.float32(1.0 / 30.0)
G = np.float32(-9.81)


def worlds(env_name, n=1, G_over=None, gravity=None, flags=None):
  env = blcd.env_map[env_name](G_over or {})
  spec = env.layout.spec
  if gravity is not None:
    spec.gravity[0], spec.gravity[1] = gravity
  if flags is not None:
    spec.flags = flags
  return env, oracle.OracleWorlds(spec, n)


def body(x, y, a=0.0, vx=0.0, vy=0.0, w=0.0):
  return [x, y, a, vx, vy, w]


def test_free_fall_is_semi_implicit_euler():
  env, ow = worlds('Bounce')
  ow.set_bodies(np.array([[body(2.5, 4.0)]], np.float32))
  ow.step(np.zeros((1, 1), np.float32))
  out = ow.get_bodies()[0, 0]
  v = np.float32(0.0); y = np.float32(4.0)
  for _ in range(3):
    v = np.float32(v + np.float32(H * G)); y = np.float32(y + np.float32(H * v))
  assert out[4] == v and out[1] == y and out[0] == np.float32(2.5) and out[3] == 0.0


def test_mass_data_box_circle_polygon():
  env, ow = worlds('Dropbox')
  ow.set_bodies(np.array([[body(2.5, 2.5)]], np.float32))
  m, I, cx, cy = ow.mass(0)
  assert m == pytest.approx(0.1 * 1.4 * 1.4, rel=1e-6) and I == pytest.approx(m * (0.7**2 + 0.7**2) / 3, rel=1e-5) and cx == 0 and cy == 0
  env, ow = worlds('Bounce')
  ow.set_bodies(np.array([[body(2.5, 2.5)]], np.float32))
  m, I, _, _ = ow.mass(0)
  assert m == pytest.approx(0.1 * math.pi * 0.25, rel=1e-6) and I == pytest.approx(0.5 * m * 0.25, rel=1e-6)
  env, ow = worlds('Luxo')
  ow.reset()
  m, I, cx, cy = ow.mass(0)   # luxo head: trapezoid, area = (40 + 24) / 2 * 28 px^2 at 30 px/m
  area = (0.8 * 50 + 0.8 * 30) / 2 * (0.8 * 35) / 900.0
  assert m == pytest.approx(0.1 * area, rel=1e-5) and abs(cy) < 1e-6 and cx > 0


def test_box_comes_to_rest_at_skin_minus_slop_and_sleeps():
  env, ow = worlds('Dropbox')
  ow.set_bodies(np.array([[body(2.5, 0.9)]], np.float32))
  for _ in range(40):
    ow.step(np.zeros((1, 1), np.float32))
  x, y, a, vx, vy, w = ow.get_bodies()[0, 0]
  # rests on the edge with both skins (2 x b2_polygonRadius) minus b2_linearSlop of allowed penetration
  assert y == pytest.approx(0.7 + 0.02 - 0.005, abs=1.5e-3) and abs(a) < 1e-3 and x == pytest.approx(2.5, abs=1e-3)
  assert vx == 0 and vy == 0 and w == 0, 'island should be asleep (velocities zeroed) after 0.5 s at rest'
  assert ow.counters()[0][oracle.COUNTER_NAMES.index('sleep_steps')] > 0


def test_no_sleep_flag_keeps_box_awake():
  env, ow = worlds('Dropbox', flags=8)
  ow.set_bodies(np.array([[body(2.5, 0.9)]], np.float32))
  for _ in range(40):
    ow.step(np.zeros((1, 1), np.float32))
  assert ow.counters()[0][oracle.COUNTER_NAMES.index('sleep_steps')] == 0


def test_ball_bounce_restitution_ratio():
  env, ow = worlds('Bounce')
  ow.set_bodies(np.array([[body(2.5, 3.0)]], np.float32))
  vy = []
  for _ in range(30):
    ow.step(np.zeros((1, 1), np.float32))
    vy.append(ow.get_bodies()[0, 0, 4])
  vy = np.array(vy)
  i = int(np.argmax(vy > 0))           # first step moving up = just after the first impact
  v_in = -math.sqrt(2 * 9.81 * (3.0 - 0.51))
  assert i > 0 and vy[i] / -v_in == pytest.approx(0.8, abs=0.08)
  assert ow.counters()[0][oracle.COUNTER_NAMES.index('toi_events')] >= 1, 'first impact must be caught by continuous collision'


def test_fast_ball_does_not_tunnel():
  env, ow = worlds('Bounce', n=4)
  b = np.array([[body(2.5, 2.5, vy=-300.0)], [body(2.5, 2.5, vx=300.0)], [body(2.5, 2.5, vx=-200.0, vy=250.0)], [body(1.0, 1.0, vx=-50.0, vy=-50.0)]], np.float32)
  ow.set_bodies(b)
  for _ in range(5):
    ow.step(np.zeros((4, 1), np.float32))
    out = ow.get_bodies()[:, 0]
    assert (out[:, 0] > 0.45).all() and (out[:, 0] < 4.55).all() and (out[:, 1] > 0.45).all() and (out[:, 1] < 4.55).all()


def test_ball_ball_collision_conserves_momentum_without_gravity():
  env, ow = worlds('Bounce2', gravity=(0.0, 0.0))
  ow.set_bodies(np.array([[body(1.5, 2.5, vx=2.0), body(3.5, 2.5, vx=-1.0)]], np.float32))
  for _ in range(6):
    ow.step(np.zeros((1, 1), np.float32))
  out = ow.get_bodies()[0]
  assert out[0, 3] + out[1, 3] == pytest.approx(1.0, abs=1e-4) and abs(out[0, 4]) < 1e-5
  # relative speed after = restitution x relative speed before (equal masses, head on)
  assert out[1, 3] - out[0, 3] == pytest.approx(0.8 * 3.0, rel=2e-2)


def test_urchin_legs_stay_within_their_limits_around_the_assembly_pose():
  """limits [-1, 1] are relative to the angle each leg is assembled at (0, 2.0, 4.2 rad from the root): pybox2d fills
  referenceAngle = bodyB.angle - bodyA.angle when the joint is defined (pinned by the recorded Urchin episode)"""
  env, ow = worlds('Urchin', n=16)
  ow.reset()
  rng = np.random.RandomState(0)
  slop = 2.0 / 180 * math.pi
  devs = []
  for t in range(40):
    ow.step(np.sign(rng.uniform(-1, 1, (16, 3))).astype(np.float32))   # full speed either way
    b = ow.get_bodies()
    for leg, rest in zip((1, 2, 3), (0.0, 2.0, 4.2)):
      rel = b[:, leg, 2] - b[:, 0, 2] - rest
      devs.append(np.abs(rel - 2 * math.pi * np.round(rel / (2 * math.pi))))
  devs = np.array(devs)
  # Box2D's limits are soft (a leg jammed against the floor can be held past its stop), but nothing like the 2 rad a
  # zero reference angle would make of the legs assembled at 2.0 and 4.2 rad
  assert (devs <= 1.0 + slop + 0.1).mean() > 0.85 and (devs <= 1.0 + slop + 0.3).mean() > 0.97 and devs.max() < 1.7
  assert devs.max() > 0.9     # and the motors do drive the legs to their stops
  # revolute anchors stay pinned: leg anchor (0, 20/30) in leg frame == root origin (to a few centimetres: continuous
  # collision moves bodies without regard for joints, and full-speed motors against the floor keep re-opening the hinge)
  gaps = []
  for leg in (1, 2, 3):
    ax = b[:, leg, 0] - np.sin(b[:, leg, 2]) * (20 / 30)
    ay = b[:, leg, 1] + np.cos(b[:, leg, 2]) * (20 / 30)
    gaps.append(np.hypot(ax - b[:, 0, 0], ay - b[:, 0, 1]))
  gaps = np.array(gaps)
  assert (gaps < 0.02).mean() > 0.9 and gaps.max() < 0.25


def test_motor_reaches_commanded_speed_in_free_space():
  env, ow = worlds('Urchin', gravity=(0.0, 0.0))
  st = np.zeros((1, 4, 6), np.float32)
  st[0, 0, :2] = (5.0, 2.5)
  for i, ang in enumerate((0.0, 2.0, 4.2)):   # legs hinged at the root centre, anchor (0, 20/30) in the leg frame
    a = math.atan2(math.sin(ang), math.cos(ang))
    st[0, i + 1, :3] = (5.0 + math.sin(a) * 20 / 30, 2.5 - math.cos(a) * 20 / 30, a)
  ow.set_bodies(st)
  ow.step(np.array([[1.0, 0.0, 0.0]], np.float32))
  b = ow.get_bodies()[0]
  # aleg starts at joint angle 0 (inside [-1, 1]) so its motor runs freely: relative angular velocity -> +8 rad/s
  assert b[1, 5] - b[0, 5] == pytest.approx(8.0, abs=0.05)
  # action is clipped to [-1, 1] (world_env.py:441)
  ow.set_bodies(st)
  ow.step(np.array([[-5.0, 0.0, 0.0]], np.float32))
  b = ow.get_bodies()[0]
  assert b[1, 5] - b[0, 5] == pytest.approx(-8.0, abs=0.05)


def test_bodies_stay_inside_walls_over_long_random_rollouts():
  for name in ('Urchin', 'LuxoCube', 'UrchinBall', 'Object2'):
    env, ow = worlds(name, n=32)
    ow.reset()
    ow.rollout(60, want=())
    b = ow.get_bodies()
    assert np.isfinite(b).all()
    W, Hh = env.WIDTH, env.HEIGHT
    assert (b[..., 0] > -0.1).all() and (b[..., 0] < W + 0.1).all() and (b[..., 1] > -0.1).all() and (b[..., 1] < Hh + 0.1).all(), name


def test_reset_distribution_matches_reference_ranges():
  # urchin root x in [1.25, 8.75], y == 1.25; Bounce2 x, y in [0.5, 4.5]; cube y in [0.4, 1.875]
  env, ow = worlds('Urchin', n=512)
  ow.reset()
  b = ow.get_bodies()
  assert b[:, 0, 0].min() >= 1.25 and b[:, 0, 0].max() <= 8.75 and (b[:, 0, 1] == np.float32(1.25)).all()
  assert b[:, 0, 0].max() - b[:, 0, 0].min() > 6.5
  env, ow = worlds('Bounce2', n=512)
  ow.reset()
  b = ow.get_bodies()
  assert b[..., :2].min() >= 0.5 and b[..., :2].max() <= 4.5
  env, ow = worlds('LuxoCube', n=512)
  ow.reset()
  b = ow.get_bodies()
  assert (b[:, 0, 1] == 2.0).all() and (b[:, 0, 2] == 0).all()
  assert b[:, 4, 1].min() >= 0.4 and b[:, 4, 1].max() <= 1.875 + 1e-6
  assert np.allclose(b[:, 1, 2], -0.5) and np.allclose(b[:, 2, 2], 0.5) and np.allclose(b[:, 3, 2], 0.0)
  assert np.allclose(b[:, 3, 1], 0.188, atol=2e-3)   # lfoot centre height quoted in SURVEY.md 8c


def test_reset_from_full_state_round_trips_through_observe():
  env, ow = worlds('UrchinBall', n=8)
  ow.reset()
  obs = ow.observe()
  ow2 = oracle.OracleWorlds(env.layout.spec, 8, seed=99)
  ow2.reset(full_state=obs['full_state'])
  obs2 = ow2.observe()
  assert np.abs(obs2['full_state'] - obs['full_state']).max() < 2e-6
  assert (obs2['lcd_bits'] == obs['lcd_bits']).mean() > 0.97
  assert obs['proprio'].shape == (8, 16) and (obs['proprio'] == obs['full_state'][:, env.pobs_idxs]).all()


def test_rollout_is_deterministic_and_independent_of_batching():
  env, a = worlds('Urchin', n=6)
  a.reset()
  ra = a.rollout(20)
  b = oracle.OracleWorlds(env.layout.spec, 3, seed=0, world_offset=3, threads=3)
  b.reset()
  rb = b.rollout(20)
  for k in ('full_state', 'lcd_bits', 'action'):
    assert (ra[k][3:] == rb[k]).all()
  assert ra['action'].min() >= -1 and ra['action'].max() <= 1 and abs(ra['action'].mean()) < 0.1
