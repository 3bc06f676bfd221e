"""Harness that runs the UNMODIFIED reference (`/root/reference/boxLCD`) under stub modules.

Only usable in the build container (where /root/reference exists); it is the generator of the
committed fixtures in this directory and is never imported by product code, bench.py or GPU tests.

What runs for real: reference `WorldEnv.__init__` (spaces / key order), `reset()` / `_reset_bodies()`
(initial-pose sampling and child placement algebra), `_get_obs()` and `lcd_render()` on the container's
Pillow.  What is stubbed: pybox2d (`Box2D`), `gym`, `pyglet` -- none of them is installed here.
The Box2D stub records bodies/joints and implements only `b2Transform * v` (fp32, separately rounded
mul/add, as the x86 build of `b2Mul(b2Transform, b2Vec2)` does).  `b2World.Step` is a no-op: the
physics itself is NOT covered by this harness; it is pinned against the recorded pybox2d episodes (tests/test_gif_hires.py,
tests/test_gif_episodes.py; oracle/README.md).
"""
import sys
import types
import math
import numpy as np

REF = '/root/reference'
F = np.float32


class _Vec2(np.ndarray):
  """np.array-like b2Vec2 (fp32 storage) with .x/.y"""
  def __new__(cls, xy):
    obj = np.asarray([F(xy[0]), F(xy[1])], dtype=np.float32).view(cls)
    return obj
  @property
  def x(self): return float(self[0])
  @property
  def y(self): return float(self[1])


class circleShape:
  def __init__(self, radius=0.0, pos=(0, 0), **kw):
    self.radius = float(F(radius))
    self.pos = (float(F(pos[0])), float(F(pos[1])))


def _hull(ps):
  """b2PolygonShape::Set gift wrapping (start right-most, lowest y on ties; CCW)."""
  n = len(ps)
  i0 = 0
  for i in range(1, n):
    if ps[i][0] > ps[i0][0] or (ps[i][0] == ps[i0][0] and ps[i][1] < ps[i0][1]):
      i0 = i
  hull = []
  ih = i0
  while True:
    hull.append(ih)
    ie = 0
    for j in range(1, n):
      if ie == ih:
        ie = j
        continue
      r = (F(ps[ie][0] - ps[ih][0]), F(ps[ie][1] - ps[ih][1]))
      v = (F(ps[j][0] - ps[ih][0]), F(ps[j][1] - ps[ih][1]))
      c = F(F(r[0] * v[1]) - F(r[1] * v[0]))
      if c < 0:
        ie = j
      if c == 0 and (v[0] * v[0] + v[1] * v[1]) > (r[0] * r[0] + r[1] * r[1]):
        ie = j
    ih = ie
    if ie == i0:
      break
  return [ps[i] for i in hull]


class polygonShape:
  def __init__(self, box=None, vertices=None, **kw):
    if box is not None:
      hx, hy = F(box[0]), F(box[1])
      vs = [(-hx, -hy), (hx, -hy), (hx, hy), (-hx, hy)]
    else:
      vs = _hull([(F(x), F(y)) for x, y in vertices])
    self.vertices = [(float(x), float(y)) for x, y in vs]
    self.radius = float(F(0.01))


class edgeShape:
  def __init__(self, vertices=None, **kw):
    self.vertices = vertices


class fixtureDef:
  def __init__(self, **kw):
    self.__dict__.update(kw)


class revoluteJointDef:
  def __init__(self, **kw):
    self.__dict__.update(kw)


class _Transform:
  def __init__(self, body):
    self._b = body
  @property
  def position(self):
    return self._b.position
  @position.setter
  def position(self, v):
    self._b.position = v
  @property
  def angle(self):
    return self._b.angle
  def __mul__(self, v):
    b = self._b
    s, c = b.sincos()
    px, py = F(b.position[0]), F(b.position[1])
    vx, vy = F(v[0]), F(v[1])
    x = F(F(F(c * vx) - F(s * vy)) + px)
    y = F(F(F(s * vx) + F(c * vy)) + py)
    return (float(x), float(y))


class FakeBody:
  """records a body; position is an fp32 pair, angle a float (stored as fp32)"""
  def __init__(self, position=(0, 0), angle=0.0, fixtures=None, sc=None, **kw):
    self._pos = _Vec2(position)
    self._angle = float(F(angle))
    kw['_angle64'] = float(angle)   # the float64 Python value before pybox2d stores it as float32
    self._sc = sc  # optional explicit (sin, cos) fp32 pair overriding sinf/cosf(angle)
    self.fixtures = fixtures if isinstance(fixtures, (list, tuple)) else [fixtures]
    self.kw = kw
    self.color1 = (0.5, 0.4, 0.9)
    self.color2 = (0.3, 0.3, 0.5)
    self.transform = _Transform(self)
  @property
  def position(self):
    return self._pos
  @position.setter
  def position(self, v):
    self._pos = _Vec2(v)
  @property
  def angle(self):
    return self._angle
  @angle.setter
  def angle(self, a):
    self._angle = float(F(a))
    self._sc = None
  def sincos(self):
    if self._sc is not None:
      return F(self._sc[0]), F(self._sc[1])
    # correctly rounded fp32 sin/cos of the fp32 angle (what glibc sinf/cosf return in all but rare cases)
    return F(math.sin(self._angle)), F(math.cos(self._angle))


class FakeJoint:
  def __init__(self, jd):
    self.jd = jd
    self.bodyA = jd.bodyA
    self.bodyB = jd.bodyB
    self.motorSpeed = jd.motorSpeed
    self.maxMotorTorque = jd.maxMotorTorque
  @property
  def angle(self):
    return self.bodyB.angle - self.bodyA.angle


class b2World:
  def __init__(self, gravity=(0, -10), doSleep=True, **kw):
    self.gravity = gravity
    self.bodies = []
    self.statics = []
    self.joints = []
    self.n_steps = 0
  def CreateStaticBody(self, **kw):
    b = FakeBody(position=(0, 0), angle=0.0, fixtures=[fixtureDef(shape=kw.get('shapes'))])
    self.statics.append(b)
    return b
  def CreateDynamicBody(self, position=(0, 0), angle=0.0, fixtures=None, **kw):
    b = FakeBody(position=position, angle=angle, fixtures=fixtures, **kw)
    self.bodies.append(b)
    return b
  def CreateJoint(self, jd):
    j = FakeJoint(jd)
    self.joints.append(j)
    return j
  def DestroyBody(self, b):
    pass
  def Step(self, dt, vi, pi):
    self.n_steps += 1


class _Box:
  def __init__(self, low, high, shape, dtype=np.float32):
    self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)
  def sample(self):
    return np.random.uniform(-1, 1, self.shape).astype(self.dtype)


class _Dict:
  def __init__(self, spaces):
    self.spaces = spaces


def install_stubs():
  if 'boxLCD' in sys.modules and getattr(sys.modules['boxLCD'], '__file__', '').startswith(REF):
    return
  if not hasattr(np, 'float'):
    np.float = float
  if not hasattr(np, 'bool'):
    np.bool = bool
  box2d = types.ModuleType('Box2D')
  box2d.b2World = b2World
  b2 = types.ModuleType('Box2D.b2')
  for k, v in dict(edgeShape=edgeShape, circleShape=circleShape, fixtureDef=fixtureDef, polygonShape=polygonShape,
                   frictionJointDef=object, contactListener=object, revoluteJointDef=revoluteJointDef).items():
    setattr(b2, k, v)
  box2d.b2 = b2
  gym = types.ModuleType('gym')
  class Env: pass
  gym.Env = Env
  spaces = types.ModuleType('gym.spaces')
  spaces.Box = _Box
  spaces.Dict = _Dict
  gym.spaces = spaces
  utils = types.ModuleType('gym.utils')
  class EzPickle:
    def __init__(self, *a, **k): pass
  utils.EzPickle = EzPickle
  seeding = types.ModuleType('gym.utils.seeding')
  def np_random(seed=None):
    rs = np.random.RandomState(seed)
    return rs, seed
  seeding.np_random = np_random
  utils.seeding = seeding
  envs = types.ModuleType('gym.envs')
  cc = types.ModuleType('gym.envs.classic_control')
  rendering = types.ModuleType('gym.envs.classic_control.rendering')
  cc.rendering = rendering
  envs.classic_control = cc
  gym.envs = envs
  gym.utils = utils
  pyglet = types.ModuleType('pyglet')
  mods = {'Box2D': box2d, 'Box2D.b2': b2, 'gym': gym, 'gym.spaces': spaces, 'gym.utils': utils, 'gym.utils.seeding': seeding,
          'gym.envs': envs, 'gym.envs.classic_control': cc, 'gym.envs.classic_control.rendering': rendering, 'pyglet': pyglet,
          'pyglet.gl': types.ModuleType('pyglet.gl')}
  sys.modules.update(mods)
  sys.path.insert(0, REF)


def ref_envs():
  install_stubs()
  import boxLCD  # noqa  (the reference, from /root/reference)
  assert boxLCD.__file__.startswith(REF), boxLCD.__file__
  return boxLCD


def render_poses(env, poses, width=None, height=None):
  """poses: list of (px, py, sin, cos) fp32 per dynamic body, in env.dynbodies order (after one reset)."""
  for (name, body), (px, py, s, c) in zip(env.dynbodies.items(), poses):
    body.position = (px, py)
    body._sc = (F(s), F(c))
  return env.lcd_render(width, height) if width else env.lcd_render()
