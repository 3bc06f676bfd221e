"""Scene-spec structs and robot fillers: the same public surface as the reference's `boxLCD/world_defs.py`
(SCALE :8, Object :11-23, Body :26-31, Joint :33-41, Robot :43-52, WorldDef :55-59, ROBOT_FILLER/register :63-70,
make_urchin :78-95, make_luxo :97-124, make_quad :129-146, make_legs :149-164).

A WorldDef built from these is pure data; `boxlcd_b200.spec.compile_spec` flattens it into the `blcd_spec` POD that
libboxlcd_b200 consumes.  The large robots (crab / walker / gingy / octo / spider, world_defs.py:168-445) are outside
round-1 scope and raise NotImplementedError when requested.
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


def _later(kind):
  def fn(robot, G):
    raise NotImplementedError(f"robot type '{kind}' (reference world_defs.py:168-445) is not built yet: round-1 scope is urchin/luxo/quad/legs")
  return fn


for _kind in ('crab', 'walker', 'gingy', 'octo', 'spider'):
  ROBOT_FILLER[_kind] = _later(_kind)
