"""`WorldEnv`: the reference's gym-style single-environment API (boxLCD/world_env.py:21-535) on top of the batched CUDA
simulator.  One `WorldEnv` owns a one-world `VecWorldEnv`; every method below is host glue around calls into
libboxlcd_b200 (reset -> blcd_reset, step -> blcd_step_host, lcd_render -> blcd_render_poses).  There is no CPU
fallback: without the CUDA library and a GPU these methods raise.

Configuration (`ENV_DG`, `G` may be a dict or a Namespace), observation / action key order, spaces and the 4-tuple
`step` return are the reference's (world_env.py:32-45, 47-142, 431-458).
"""
import numpy as np
from boxlcd_b200 import utils, spaces
from boxlcd_b200.world_defs import SCALE, ROBOT_FILLER
from boxlcd_b200.spec import compile_spec

A = utils.A


class WorldEnv:
  metadata = {'render.modes': ['human', 'rgb_array']}
  # ENVIRONMENT DEFAULT CONFIG (world_env.py:32-45)
  ENV_DG = utils.AttrDict(
      base_dim=5,        # base size of the physics world (HEIGHT)
      lcd_base=16,       # frame height in pixels; width = wh_ratio * height
      wh_ratio=2.0,      # width:height ratio of world and frames
      ep_len=100,        # steps before done/timeout
      angular_offset=0, root_offset=0, compact_obs=0,   # alternative observation layouts (not built)
      use_speed=1,       # velocity control (torque control is broken in the reference)
      all_corners=0, walls=1, debug=0, fps=10,
  )

  def __init__(self, world_def, G={}, device=None, seed=None):
    self.world_def = world_def
    self.G = utils.AttrDict(self.ENV_DG)
    if not isinstance(G, dict):
      G = G.__dict__
    for key in G:
      self.G[key] = G[key]
    self.scroll = 0.0
    self.viewer = None
    # fill partial robot descriptions in place, like the reference (world_env.py:83-85)
    for i, robot in enumerate(self.world_def.robots):
      if robot.root_body is None:
        self.world_def.robots[i] = ROBOT_FILLER[robot.type](robot, self.G)
    self.layout = compile_spec(self.world_def, self.G, self.WIDTH, self.HEIGHT)
    self.obs_info, self.act_info = self.layout.obs_info, self.layout.act_info
    self.obs_size = len(self.obs_info)
    self.obs_keys = list(self.obs_info.keys())
    self.pobs_keys = utils.nfiltlist(self.obs_keys, 'object')
    self.pobs_size = len(self.pobs_keys)
    self.pobs_idxs = [self.obs_keys.index(x) for x in self.pobs_keys]
    sp = {}
    sp['full_state'] = spaces.Box(-1, +1, (self.obs_size,), dtype=np.float32)
    sp['proprio'] = spaces.Box(-1, +1, (max(self.pobs_size, 1),), dtype=np.float32)
    sp['lcd'] = spaces.Box(0, 1, (self.G.lcd_base, int(self.G.lcd_base * self.G.wh_ratio)), dtype=np.bool_)
    self.observation_space = spaces.Dict(sp)
    self.act_size = len(self.act_info)
    self.act_keys = list(self.act_info.keys())
    self.action_space = spaces.Box(-1, +1, (self.act_size,), dtype=np.float32)
    self._device = device
    self._vec = None
    self._seed = seed
    self.ep_t = 0
    self.seed(seed)

  # -- geometry properties (world_env.py:144-166) ---------------------------------------------------------------------
  @property
  def WIDTH(self): return int(self.G.wh_ratio * self.G.base_dim)
  @property
  def HEIGHT(self): return self.G.base_dim
  @property
  def VIEWPORT_H(self): return 30 * self.HEIGHT
  @property
  def VIEWPORT_W(self): return 30 * self.WIDTH
  @property
  def FPS(self): return self.G.fps
  @property
  def SCALE(self): return SCALE

  def seed(self, seed=None):
    """world_env.py:168-170.  The per-world Philox stream of the simulator is re-keyed with `seed`."""
    if seed is None:
      seed = int(np.random.SeedSequence().entropy % (2**63))
    self._seed = seed
    if self._vec is not None:
      self._vec.close()
      self._vec = None
    return [seed]

  def _sim(self):
    if self._vec is None:
      from boxlcd_b200.vec_env import VecWorldEnv
      self._vec = VecWorldEnv(self, 1, device=self._device, seed=self._seed)
    return self._vec

  def close(self):
    if self.viewer is not None:
      self.viewer.window.close()
      self.viewer = None
    if self._vec is not None:
      self._vec.close()
      self._vec = None

  # -- gym API --------------------------------------------------------------------------------------------------------
  def reset(self, full_state=None, proprio=None):
    """world_env.py:306-385"""
    self.ep_t = 0
    if proprio is not None:
      proprio = np.asarray(proprio)
      assert proprio.shape[-1] == self.observation_space.spaces['proprio'].shape[-1], f'invalid shape for proprio {proprio.shape} {self.observation_space.spaces["proprio"]}'
      full_state = np.zeros(self.observation_space.spaces['full_state'].shape)
      full_state[self.pobs_idxs] = proprio
    vec = self._sim()
    obs = vec.reset(full_state=None if full_state is None else np.asarray(full_state, np.float32)[None])
    return self._unbatch(self._follow(obs))

  def step(self, action):
    """world_env.py:431-458"""
    self.ep_t += 1
    obs, rew, done, infos = self._sim().step(np.asarray(action, np.float32).reshape(1, self.act_size))
    done = bool(done[0])
    return self._unbatch(self._follow(obs)), 0.0, done, {'timeout': done}

  def _follow(self, obs):
    """walls=0: the view offset tracks the first robot's root (world_env.py:381-382, 453-454; frames ignore it, :462)"""
    if not self.G.walls:
      robot = self.world_def.robots[0]
      x01 = float(obs['full_state'][0, self.obs_keys.index(f'{robot.type}0:root:x:p')])   # normalized to [-1, 1] over [0, WIDTH]
      self.scroll = (x01 + 1.0) * 0.5 * self.WIDTH - self.VIEWPORT_W / SCALE / 2
    return obs

  def _get_obs(self):
    return self._unbatch(self._sim().observe())

  def _unbatch(self, obs):
    return {'full_state': obs['full_state'][0].astype(np.float64), 'proprio': obs['proprio'][0].astype(np.float64), 'lcd': obs['lcd'][0]}

  def lcd_render(self, width=None, height=None, lcd_mode='1'):
    """world_env.py:460-512.  Mode '1' at any size is rendered by the CUDA rasterizer; mode 'RGB' (the viewer's colour
    picture, world_env.py:481-483,509-511) is drawn on the host with Pillow from the simulator's body transforms
    (boxlcd_b200/rgb_render.py)."""
    lcd_mode = lcd_mode.upper()
    assert lcd_mode in ['1', 'RGB'], 'lcd_mode must be in one of these PIL supported modes'
    if width is None and height is None:
      width = int(self.G.lcd_base * self.G.wh_ratio)
      height = self.G.lcd_base
    if lcd_mode == 'RGB':
      return self._sim().render_rgb(width, height)[0]
    return self._sim().render(width, height)[0]

  def render(self, mode='rgb_array', lcd_mode='1', return_pyglet_view=False):
    """world_env.py:514-535.  mode='human' composes [8x colour view | 1 px | LCD frame x8] (world_env.py:525-531).  With
    `return_pyglet_view` the composed picture itself is returned (the reference reads the same pixels back from the
    window's colour buffer, viewer.py:31-36); showing it in a window needs pyglet, which is imported only then."""
    from boxlcd_b200 import rgb_render
    lcd_mode = lcd_mode.upper()
    width = int(self.G.lcd_base * self.G.wh_ratio)
    height = self.G.lcd_base
    lcd = self.lcd_render(width, height, lcd_mode=lcd_mode)
    if mode == 'rgb_array':
      return lcd
    elif mode == 'human':
      high_res = self.lcd_render(width * 8, height * 8, lcd_mode='RGB').astype(np.uint8)
      img = rgb_render.human_frame(high_res, lcd)
      try:
        import pyglet  # noqa: F401
      except ImportError:
        if return_pyglet_view:
          return img
        raise ImportError("render(mode='human') shows a window and needs pyglet; pass return_pyglet_view=True to get the composed picture instead")
      if self.viewer is None:
        self.viewer = _Viewer(width * 8, height * 8)
      self.viewer.render(img)
      return img if return_pyglet_view else lcd


class _Viewer:
  """viewer.py:4-37: a pyglet window that blits an already rendered picture"""
  def __init__(self, width, height):
    import pyglet
    self.pyglet = pyglet
    self.window = pyglet.window.Window(2 * width, height)

  def render(self, image):
    self.window.clear()
    self.window.switch_to()
    self.window.dispatch_events()
    self.pyglet.image.ImageData(image.shape[1], image.shape[0], 'RGB', image.tobytes(), pitch=image.shape[1] * -3).blit(0, 0)
    self.window.flip()
