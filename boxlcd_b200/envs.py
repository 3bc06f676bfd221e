"""Named environments = scene spec + per-env config defaults, the same public surface as the reference's
`boxLCD/envs.py` (cc :5-14; Dropbox :17, Bounce :23, Bounce2 :29, Object2 :35, Object3 :41, Urchin :48, Luxo :54,
UrchinCube :66, LuxoCube :72, UrchinBall :79, LuxoBall :86, *Balls / *Cubes :92-110, Crab / CrabCube / SpiderCube :116-137)."""
from boxlcd_b200.world_env import WorldEnv
from boxlcd_b200.world_defs import WorldDef, Object, Robot
from boxlcd_b200 import utils


def cc(**overrides):
  """class decorator: subclass whose ENV_DG is WorldEnv.ENV_DG with `overrides` applied (envs.py:5-14)"""
  def decorator(Cls):
    class CustomWorldEnv(Cls):
      ENV_DG = utils.AttrDict(WorldEnv.ENV_DG, **overrides)
    CustomWorldEnv.__name__ = Cls.__name__
    CustomWorldEnv.__qualname__ = Cls.__qualname__
    return CustomWorldEnv
  return decorator


def _env(name, robots=(), objects=(), **overrides):
  """build a named env class from robot types and object specs"""
  def __init__(self, G={}, **kw):
    w = WorldDef(robots=[Robot(type=t, name=f'{t}0') for t in robots], objects=[Object(f'object{i}', **o) for i, o in enumerate(objects)])
    WorldEnv.__init__(self, w, G, **kw)
  cls = type(name, (WorldEnv,), {'__init__': __init__, '__module__': __name__})
  return cc(**overrides)(cls) if overrides else cls


cube_settings = dict(shape='box', size=0.4, density=0.5, linearDamping=1.0, angularDamping=0.2)
ball_settings = dict(shape='circle', size=0.5, density=0.2, restitution=0.8)
_bouncy = dict(size=0.5, density=0.1, restitution=0.8)

# BASIC PASSIVE ENVS
Dropbox = _env('Dropbox', objects=[dict(shape='box', size=0.7, density=0.1)], ep_len=25, wh_ratio=1.0)
Bounce = _env('Bounce', objects=[dict(shape='circle', **_bouncy)], ep_len=50, wh_ratio=1.0)
Bounce2 = _env('Bounce2', objects=[dict(shape='circle', **_bouncy)] * 2, ep_len=50, wh_ratio=1.0)
Object2 = _env('Object2', objects=[dict(shape='random', **_bouncy)] * 2, ep_len=50, wh_ratio=1.0)
Object3 = _env('Object3', objects=[dict(shape='random', **_bouncy)] * 3, ep_len=50, wh_ratio=1.0)
# SIMPLE ROBOTS
Urchin = _env('Urchin', robots=['urchin'], ep_len=100)
Luxo = _env('Luxo', robots=['luxo'], ep_len=100)
# SIMPLE ROBOT OBJECT MANIPULATION
UrchinCube = _env('UrchinCube', robots=['urchin'], objects=[cube_settings], ep_len=150, wh_ratio=1.5)
LuxoCube = _env('LuxoCube', robots=['luxo'], objects=[cube_settings], ep_len=150, wh_ratio=1.5)
UrchinBall = _env('UrchinBall', robots=['urchin'], objects=[ball_settings], ep_len=150, wh_ratio=1.5)
LuxoBall = _env('LuxoBall', robots=['luxo'], objects=[ball_settings], ep_len=150, wh_ratio=1.5)
UrchinBalls = _env('UrchinBalls', robots=['urchin'], objects=[ball_settings] * 3)
LuxoBalls = _env('LuxoBalls', robots=['luxo'], objects=[ball_settings] * 3)
UrchinCubes = _env('UrchinCubes', robots=['urchin'], objects=[cube_settings] * 3)
LuxoCubes = _env('LuxoCubes', robots=['luxo'], objects=[cube_settings] * 3)
# MORE ADVANCED: 64 x 32 frames, up to 18 bodies -- the library's large-scene profile
Crab = _env('Crab', robots=['crab'], lcd_base=32)
CrabCube = _env('CrabCube', robots=['crab'], objects=[dict(shape='box', size=0.4, density=1.0, friction=1.0)], lcd_base=32)
SpiderCube = _env('SpiderCube', robots=['spider'], objects=[dict(shape='box', size=0.3, density=0.1, friction=1.0)], lcd_base=32)
