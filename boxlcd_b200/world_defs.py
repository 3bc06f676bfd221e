"""Scene-spec structs and robot fillers: the same public surface as the reference's `boxLCD/world_defs.py`
(SCALE :8, Object :11-23, Body :26-31, Joint :33-41, Robot :43-52, WorldDef :55-59, ROBOT_FILLER/register :63-70,
make_urchin :78-95, make_luxo :97-124, make_quad :129-146, make_legs :149-164, crab/walker/gingy/octo/spider :168-445).

A WorldDef built from these is pure data; `boxlcd_b200.spec.compile_spec` flattens it into the `blcd_spec` POD that
libboxlcd_b200 consumes.  The large robots (crab / walker / gingy / octo / spider, world_defs.py:168-445) are table-driven (`_limbed`); scenes
with more than 8 bodies run on the library's large-scene profile.
"""
from typing import NamedTuple, List, Tuple, Dict, Any
from boxlcd_b200.shapes import circleShape, polygonShape

SCALE = 30.0  # pixels per metre of the 30 px/m design grid every robot dimension below is written in


class Object(NamedTuple):
  name: str
  shape: str = 'box'          # 'box' | 'circle' | 'random'
  size: float = 0.5           # half extent / radius
  linearDamping: float = 0.0
  angularDamping: float = 0.0
  density: float = 1.0
  friction: float = 0.5
  restitution: float = 0.0
  categoryBits: int = 0x0110
  rand_angle: int = 1
  rangex: Tuple[float, float] = None
  rangey: Tuple[float, float] = None


class Body(NamedTuple):
  shape: Any
  density: float = 1
  maskBits: int = 0x001
  categoryBits: int = 0x0020
  friction: float = 1.0


class Joint(NamedTuple):
  parent: str
  angle: float
  anchorA: list
  anchorB: list
  limits: List[float]
  limited: bool = True
  speed: float = 8
  torque: float = 150


class Robot(NamedTuple):
  type: str
  name: str
  root_body: Body = None
  bodies: Dict[str, Body] = None
  joints: Dict[str, Joint] = None
  rand_angle: int = 0
  angularDamping: float = 0
  linearDamping: float = 0
  bound: float = 1.5


class WorldDef(NamedTuple):
  robots: List[Robot] = []
  objects: List[Object] = []
  gravity: List[float] = [0, -9.81]
  forcetorque: int = 0


ROBOT_FILLER = {}


def register(name):
  def deco(fn):
    ROBOT_FILLER[name] = fn
    return fn
  return deco


def _px(*v):
  """design-grid pixels -> metres"""
  return tuple(x / SCALE for x in v) if len(v) > 1 else v[0] / SCALE


def _star(robot, leg_angles, rand_angle, bound):
  """circle hub with identical box legs hinged at the hub centre (urchin / quad / legs)"""
  leg_w, leg_h = _px(8), _px(40)
  leg = polygonShape(box=(leg_w / 2, leg_h / 2))
  names = [f'{c}leg' for c in 'abcdefgh'[:len(leg_angles)]]
  return Robot(type=robot.type, name=robot.name,
               root_body=Body(circleShape(radius=0.8 * leg_w)),
               bodies={n: Body(leg, maskBits=0x011, density=1.0) for n in names},
               joints={n: Joint('root', a, (0, 0), (0, leg_h / 2), [-1.0, 1.0], limited=True) for n, a in zip(names, leg_angles)},
               rand_angle=rand_angle, bound=bound)


@register('urchin')
def make_urchin(robot, G):
  return _star(robot, (0.0, 2.0, 4.2), rand_angle=1, bound=1.25)


@register('quad')
def make_quad(robot, G):
  return _star(robot, (0.0, 2.0, 4.2), rand_angle=0, bound=1.5)


@register('legs')
def make_legs(robot, G):
  return _star(robot, (-1.0, 1.0), rand_angle=0, bound=1.5)


@register('luxo')
def make_luxo(robot, G):
  vert, side = _px(10), _px(5)
  leg_w, leg_h, shin_h = _px(8), _px(24), _px(20)
  head = [((x * 0.8) / SCALE, (y * 0.8) / SCALE) for x, y in ((-15, 15), (20, 25), (20, -25), (-15, -15))]
  parts = {
      'lhip': (polygonShape(box=(leg_w / 2, leg_h / 2)), Joint('root', -0.5, (-side, -vert), (0, leg_h / 2), [-0.1, 0.1])),
      'lknee': (polygonShape(box=(0.8 * leg_w / 2, shin_h / 2)), Joint('lhip', 0.5, (0, -leg_h / 2), (0, shin_h / 2), [-0.9, 0.9])),
      'lfoot': (polygonShape(box=(leg_h, leg_w / 2)), Joint('lknee', 0.0, (0, -leg_h / 2), (0, leg_w / 2), [-0.5, 0.9])),
  }
  return Robot(type=robot.type, name=robot.name,
               root_body=Body(polygonShape(vertices=head), density=0.1, maskBits=0x011),
               bodies={n: Body(s, maskBits=0x011) for n, (s, _) in parts.items()},
               joints={n: j for n, (_, j) in parts.items()},
               bound=2.0)


def _box(w, h):
  return polygonShape(box=(w / 2, h / 2))


def _limbed(robot, root_body, parts, **robot_kw):
  """robot from a parts table: name -> (shape, body options, parent, rest angle, anchor on parent, anchor on part, limits, joint options).
  Dict order is creation order (= draw order, world_env.py:226-268)."""
  return Robot(type=robot.type, name=robot.name, root_body=root_body,
               bodies={n: Body(p[0], **p[1]) for n, p in parts.items()},
               joints={n: Joint(p[2], p[3], p[4], p[5], list(p[6]), **p[7]) for n, p in parts.items()},
               **robot_kw)


_FREE = dict(limited=False)
_GRIP = dict(maskBits=0x011)   # also collides with category 0x0010


def _claw_pair(parts, arm, prefix, arm_h, claw, claw_h, body_kw):
  """two-finger gripper on the end of `arm`: each finger is two segments, the outer one welded by a [0, 0] limit"""
  for side, sgn, lim in (('l', 1.0, (-2.0, 1.0)), ('r', -1.0, (-1.0, 2.0))):
    parts[f'{prefix}{side}claw0'] = (claw, body_kw, arm, sgn * 2.25, (0, arm_h / 2), (0, -claw_h / 2), lim, {})
    parts[f'{prefix}{side}claw1'] = (claw, body_kw, f'{prefix}{side}claw0', sgn * 3.75, (0, claw_h / 2), (0, -claw_h / 2), (0.0, 0.0), {})


@register('crab')
def make_crab(robot, G):
  """six-sided shell, two 2-segment legs, two 2-segment arms each ending in a two-finger claw: 17 bodies, 16 hinges,
  12 of them driven (world_defs.py:168-249)"""
  vert, side = _px(12), _px(20)
  leg_w, leg_h, shin_h = _px(8), _px(20), _px(20)
  arm_w, arm_h = _px(8), _px(20)
  claw_w, claw_h = _px(4), _px(16)
  shell = [((0.9 * x) / SCALE, (0.9 * y) / SCALE) for x, y in ((-25, 0), (-20, 16), (20, 16), (25, 0), (20, -16), (-20, -16))]
  hip, knee, arm, claw = _box(leg_w, leg_h), _box(0.8 * leg_w, shin_h), _box(arm_w, arm_h), _box(claw_w, claw_h)
  parts = {
      'lhip': (hip, {}, 'root', -0.5, (-side, -vert), (0, leg_h / 2), (-1.5, 0.5), {}),
      'rhip': (hip, {}, 'root', 0.5, (side, -vert), (0, leg_h / 2), (0.5, 1.5), {}),
      'lknee': (knee, {}, 'lhip', 0.5, (0, -leg_h / 2), (0, shin_h / 2), (-0.5, 0.5), {}),
      'rknee': (knee, {}, 'rhip', -0.5, (0, -leg_h / 2), (0, shin_h / 2), (-0.5, 0.5), {}),
      'lshoulder': (arm, _GRIP, 'root', 2.0, (-side, vert), (0, -arm_h / 2), (-3.0, 3.0), _FREE),
      'rshoulder': (arm, _GRIP, 'root', -2.0, (side, vert), (0, -arm_h / 2), (-3.0, 3.0), _FREE),
      'lelbow': (arm, _GRIP, 'lshoulder', 3.0, (0, arm_h / 2), (0, -arm_h / 2), (-2.0, 2.0), _FREE),
      'relbow': (arm, _GRIP, 'rshoulder', -3.0, (0, arm_h / 2), (0, -arm_h / 2), (-2.0, 2.0), _FREE),
  }
  _claw_pair(parts, 'lelbow', 'l', arm_h, claw, claw_h, _GRIP)
  _claw_pair(parts, 'relbow', 'r', arm_h, claw, claw_h, _GRIP)
  return _limbed(robot, Body(polygonShape(vertices=shell), density=1.0), parts, bound=2.0)


@register('walker')
def make_walker(robot, G):
  """hull on two 2-segment legs with one arm + claw on top (world_defs.py:251-298)"""
  leg_down = -_px(6)
  leg_w, leg_h = _px(10), _px(24)
  arm_w, arm_h = _px(8), _px(20)
  claw_w, claw_h = _px(6), _px(16)
  hull = [((0.8 * x) / SCALE, (0.8 * y) / SCALE) for x, y in ((-30, 9), (6, 9), (34, 1), (34, -8), (-30, -8))]
  hip, knee, arm, claw = _box(leg_w, leg_h), _box(0.8 * leg_w, leg_h), _box(arm_w, arm_h), _box(claw_w, claw_h)
  light = dict(density=0.1)
  parts = {
      'lhip': (hip, {}, 'root', 0.05, (0.0, leg_down), (0, leg_h / 2), (-0.8, 1.1), {}),
      'lknee': (knee, {}, 'lhip', 0.05, (0, -leg_h / 2), (0, leg_h / 2), (-1.6, -0.1), {}),
      'rhip': (hip, {}, 'root', -0.05, (0.0, leg_down), (0, leg_h / 2), (-0.8, 1.1), {}),
      'rknee': (knee, {}, 'rhip', -0.05, (0, -leg_h / 2), (0, leg_h / 2), (-1.6, -0.1), {}),
      'shoulder': (arm, light, 'root', 2.0, (0, _px(5)), (0, -arm_h / 2), (-3.0, 3.0), _FREE),
      'elbow': (arm, light, 'shoulder', 3.0, (0, arm_h / 2), (0, -arm_h / 2), (-2.0, 2.0), _FREE),
  }
  _claw_pair(parts, 'elbow', '', arm_h, claw, claw_h, dict(maskBits=0x011, density=0.1))
  return _limbed(robot, Body(polygonShape(vertices=hull)), parts)


@register('gingy')
def make_gingy(robot, G):
  """gingerbread figure: light round head on a torso with two 2-segment arms and two legs (world_defs.py:301-335)"""
  vert, side = _px(10), _px(2)
  body_w, body_h = _px(8), _px(25)
  arm_w, arm_h = _px(8), _px(25)
  leg_w, leg_h = _px(8), _px(30)
  torso, arm, leg = _box(body_w, body_h), _box(arm_w, arm_h), _box(leg_w, leg_h)
  heavy = dict(density=1.0)
  parts = {
      'body': (torso, heavy, 'root', 0.0, (0, -vert), (0, body_h / 2), (-0.1, 0.1), {}),
      'larm': (arm, _GRIP, 'body', 1.5, (-side, +vert), (0, arm_h / 2), (-1.5, 0.8), {}),
      'rarm': (arm, _GRIP, 'body', -1.5, (side, +vert), (0, arm_h / 2), (-1.5, 0.8), {}),
      'llarm': (arm, _GRIP, 'larm', 1.5, (0, -arm_h / 2), (0, arm_h / 2), (-1.5, 1.5), {}),
      'rlarm': (arm, _GRIP, 'rarm', -1.5, (0, -arm_h / 2), (0, arm_h / 2), (-1.5, 1.5), {}),
      'lleg': (leg, heavy, 'body', 0.8, (-side, -vert), (0, leg_h / 2), (-0.2, 0.4), {}),
      'rleg': (leg, heavy, 'body', -0.8, (side, -vert), (0, leg_h / 2), (-0.4, 0.2), {}),
  }
  return _limbed(robot, Body(circleShape(radius=_px(10)), density=0.01), parts)


@register('octo')
def make_octo(robot, G):
  """round hub with four free-swinging 2-segment tentacles (world_defs.py:337-367)"""
  leg_w, leg_h = _px(8), _px(25)
  leg = _box(leg_w, leg_h)
  body_kw = dict(maskBits=0x011, density=1.0)
  parts = {}
  for k, c in enumerate('abcd'):
    parts[f'{c}leg1'] = (leg, body_kw, 'root', float(k), (0, 0), (0, leg_h / 2), (-1.0, 1.0), _FREE)
  for k, c in enumerate('abcd'):
    parts[f'{c}leg2'] = (leg, body_kw, f'{c}leg1', float(k), (0, -leg_h / 2), (0, leg_h / 2), (-1.0, 1.0), _FREE)
  return _limbed(robot, Body(circleShape(radius=1.5 * leg_w), density=0.1), parts, rand_angle=1)


@register('spider')
def make_spider(robot, G):
  """round hub with four 2-segment legs, two below and two (light, gripping) above (world_defs.py:370-445; the arm and
  claw bodies the reference lists there have no joints and are therefore never created, world_env.py:226)"""
  vert, side = _px(8), _px(8)
  leg_w, leg_h, shin_h = _px(6), _px(20), _px(20)
  arm_w, arm_h = _px(6), _px(26)
  hip, knee, arm = _box(leg_w, leg_h), _box(0.8 * leg_w, shin_h), _box(arm_w, arm_h)
  upper = dict(maskBits=0x011, density=0.1)
  parts = {
      'lhip': (hip, {}, 'root', -1.0, (-side, -vert), (0, leg_h / 2), (-1.5, 0.5), {}),
      'rhip': (hip, {}, 'root', 1.0, (side, -vert), (0, leg_h / 2), (0.5, 1.5), {}),
      'lknee': (knee, {}, 'lhip', 0.5, (0, -leg_h / 2), (0, shin_h / 2), (-0.5, 0.5), {}),
      'rknee': (knee, {}, 'rhip', -0.5, (0, -leg_h / 2), (0, shin_h / 2), (-0.5, 0.5), {}),
      'ulhip': (arm, upper, 'root', 1.5, (-side, vert), (0, -leg_h / 2), (-1.5, 0.5), {}),
      'urhip': (arm, upper, 'root', -1.5, (side, vert), (0, -leg_h / 2), (0.5, 1.5), {}),
      'ulknee': (arm, upper, 'ulhip', -0.5, (0, leg_h / 2), (0, shin_h / 2), (-0.5, 0.5), {}),
      'urknee': (arm, upper, 'urhip', 0.5, (0, leg_h / 2), (0, shin_h / 2), (-0.5, 0.5), {}),
  }
  return _limbed(robot, Body(circleShape(radius=_px(10)), density=1.0, maskBits=0x011), parts, bound=1.3)
