"""boxlcd_b200: B200-native drop-in for boxLCD's hot path (batched Box2D-style world stepping + LCD rasterization).
Public surface mirrors `boxLCD/__init__.py:9-17`: `envs`, `env_map`, `ENV_DG`, `WorldEnv`, `WorldDef`, `Object`, `Robot`."""
import inspect
from boxlcd_b200.world_env import WorldEnv
from boxlcd_b200.world_defs import WorldDef, Object, Robot
from boxlcd_b200 import envs
from boxlcd_b200.utils import AttrDict

__version__ = '0.1.0'
ENV_DG = AttrDict(WorldEnv.ENV_DG)
env_map = {name: obj for name, obj in inspect.getmembers(envs) if inspect.isclass(obj) and issubclass(obj, WorldEnv) and obj is not WorldEnv}
