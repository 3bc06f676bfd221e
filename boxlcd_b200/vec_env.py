"""`VecWorldEnv`: N worlds of one scene, resident on one GPU, behind the call shape the reference's vector wrapper
exposes (research/wrappers/async_vector_env.py:131-242: `reset(idxs, **kwargs)`, `step(actions)` -> (dict of [N, ...]
arrays, rew[N], done[N], infos), `action_space.sample()`), plus device-tensor entry points for callers that keep their
data on the GPU.  PyTorch only owns the tensors and the stream; every operation is a call into libboxlcd_b200
(include/boxlcd_b200.h).  No CPU fallback."""
import ctypes as C
import numpy as np
import torch
from boxlcd_b200 import _lib, spaces

COUNTER_NAMES = ['contacts', 'pos_iters', 'toi_events', 'toi_calls', 'sleep_steps', 'overflow', 'manifold_points', 'substeps']


def _ptr(t):
  return None if t is None else C.c_void_p(t.data_ptr())


class VecWorldEnv:
  def __init__(self, env, num_envs, device=None, seed=0, world_offset=0):
    """env: a boxlcd_b200 WorldEnv (scene + config); num_envs worlds are created on `device` (default cuda:current)."""
    if not torch.cuda.is_available():
      raise RuntimeError('boxlcd_b200 needs a CUDA device: the simulator only exists as sm_100a kernels (no CPU fallback)')
    self.env = env
    self.spec = env.layout.spec
    self.num_envs = self.n = int(num_envs)
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    if self.device.index is None:
      self.device = torch.device('cuda', torch.cuda.current_device())
    self.l = _lib.lib()
    self.world_offset = int(world_offset)
    self.h = None
    self._create(seed)
    sp = self.spec
    self.B, self.S, self.A, self.P = sp.n_bodies, sp.obs_size, sp.act_size, max(sp.pobs_size, 1)
    self.H, self.W = sp.lcd_h, sp.lcd_w
    self.observation_space = env.observation_space
    self.single_action_space = env.action_space
    self.action_space = spaces.Box(-1, 1, (self.n, self.A), dtype=np.float32)
    f32 = dict(dtype=torch.float32, device=self.device)
    self._fs = torch.empty((self.n, self.S), **f32)
    self._pr = torch.empty((self.n, self.P), **f32)
    self._bits = torch.empty((self.n,) + self.bits_shape(), dtype=torch.int32, device=self.device)
    self._done = torch.empty((self.n,), dtype=torch.uint8, device=self.device)
    self._act = torch.empty((self.n, self.A), **f32)

  def _create(self, seed):
    h = C.c_void_p()
    _lib.check(self.l.blcd_create(C.byref(self.spec), self.n, self.device.index, int(seed or 0), self.world_offset, C.byref(h)))
    self.h, self._seed = h, int(seed or 0)

  def seed(self, seeds=None):
    """AsyncVectorEnv.seed (async_vector_env.py:108-129): re-key the per-world random streams.  Worlds draw from Philox streams
    keyed by (seed, global world index), so one integer seeds them all; a list is reduced to its first entry.  The
    simulation state is rebuilt: call reset() afterwards, as with the reference (seeding only affects later resets)."""
    if seeds is not None and not isinstance(seeds, (int, np.integer)):
      seeds = list(seeds)[0]
    self.close()
    self._create(0 if seeds is None else int(seeds))

  def rekey(self, seed=None, world_offset=None):
    """blcd_rekey: keep every allocation, make world w the global world `world_offset + w` of stream family `seed`; the next
    reset samples from those streams (as a freshly created VecWorldEnv with the same arguments would)."""
    seed = self._seed if seed is None else int(seed)
    world_offset = self.world_offset if world_offset is None else int(world_offset)
    torch.cuda.current_stream(self.device).synchronize()
    _lib.check(self.l.blcd_rekey(self.h, seed, world_offset))
    self._seed, self.world_offset = seed, world_offset

  # -- lifetime -------------------------------------------------------------------------------------------------------
  def close(self):
    if getattr(self, 'h', None):
      self.l.blcd_destroy(self.h)
      self.h = None

  __del__ = close

  def _stream(self):
    return torch.cuda.current_stream(self.device).cuda_stream

  def info(self):
    out = (C.c_int32 * 16)()
    _lib.check(self.l.blcd_scene_info(self.h, out))
    keys = ['n_bodies', 'n_joints', 'n_walls', 'n_pairs', 'obs_size', 'pobs_size', 'act_size', 'lcd_w', 'lcd_h', 'manifold_slots',
            'state_words', 'smem_words_per_world', 'block', 'smem_bytes_per_block', 'profile', 'pipeline']
    return dict(zip(keys, list(out)))

  @property
  def kernel_launches(self):
    return int(self.l.blcd_kernel_launches(self.h))

  # -- device-tensor API ------------------------------------------------------------------------------------------------
  def reset_dev(self, idx=None, full_state=None):
    """idx: int64 device tensor of world indices (None = all); full_state: [len(idx) or N, S] float32 device tensor"""
    n = 0 if idx is None else idx.numel()
    if idx is not None and n == 0:
      return   # an empty index list resets nothing (a NULL index pointer would mean "all worlds" at the C ABI)
    _lib.check(self.l.blcd_reset(self.h, _ptr(idx), n, _ptr(full_state), self._stream()))

  def step_dev(self, actions=None, observe=True):
    """actions: [N, A] float32 device tensor, or None to draw U[-1,1) actions from the per-world device RNG.
    Returns (obs dict of device tensors or None, actions used)."""
    if observe:
      _lib.check(self.l.blcd_step_observe(self.h, _ptr(actions), _ptr(self._act), _ptr(self._fs), _ptr(self._pr), _ptr(self._bits), None,
                                          _ptr(self._done), self._stream()))
      return {'full_state': self._fs, 'proprio': self._pr, 'lcd_bits': self._bits, 'done': self._done}, self._act
    _lib.check(self.l.blcd_step(self.h, _ptr(actions), _ptr(self._act), self._stream()))
    return None, self._act

  def observe_dev(self, lcd_bool=None):
    _lib.check(self.l.blcd_observe(self.h, _ptr(self._fs), _ptr(self._pr), _ptr(self._bits), _ptr(lcd_bool), _ptr(self._done), self._stream()))
    return {'full_state': self._fs, 'proprio': self._pr, 'lcd_bits': self._bits, 'done': self._done}

  def rollout_dev(self, T, full_state=None, lcd_bits=None, actions=None):
    """collect.py's inner loop for all worlds, T steps in one launch.  Output tensors [N, T, ...] are allocated if not given."""
    f32 = dict(dtype=torch.float32, device=self.device)
    full_state = torch.empty((self.n, T, self.S), **f32) if full_state is None else full_state
    lcd_bits = torch.empty((self.n, T) + self.bits_shape(), dtype=torch.int32, device=self.device) if lcd_bits is None else lcd_bits
    actions = torch.empty((self.n, T, self.A), **f32) if actions is None else actions
    _lib.check(self.l.blcd_rollout(self.h, T, _ptr(full_state), _ptr(lcd_bits), _ptr(actions), self._stream()))
    return {'full_state': full_state, 'lcd_bits': lcd_bits, 'action': actions}

  def render_poses_dev(self, poses, variants=None, width=0, height=0):
    """poses [n, B, 4] float32 (x, y, sin, cos) -> packed frames [n, H] int32 ([n, H, 2] for frames wider than 32 px)"""
    n = poses.shape[0]
    out = torch.empty((n,) + self.bits_shape(height, width), dtype=torch.int32, device=self.device)
    _lib.check(self.l.blcd_render_poses_sized(self.h, _ptr(poses), _ptr(variants), n, width, height, _ptr(out), self._stream()))
    return out

  # -- host-buffer stepping (the C ABI's reference-facing entry points) ------------------------------------------------
  def pin_host(self, *arrays):
    """page-lock numpy arrays once so that step_host_async can copy straight from / into them (or slices of them)"""
    for a in arrays:
      _lib.check(self.l.blcd_pin_host(self.h, a.ctypes.data, a.nbytes))

  def step_host(self, actions, full_state=None, lcd_bits=None, done=None):
    """one env step: numpy actions [N, A] f32 in, numpy full_state [N, S] f32 / packed frames / done [N] u8 out (synchronous)"""
    p = lambda a: None if a is None else a.ctypes.data
    _lib.check(self.l.blcd_step_host(self.h, p(actions), p(full_state), p(lcd_bits), p(done)))

  def step_host_async(self, actions, full_state=None, lcd_bits=None, done=None):
    """AsyncVectorEnv.step_async with host buffers (all inside memory given to pin_host); returns immediately"""
    p = lambda a: None if a is None else a.ctypes.data
    _lib.check(self.l.blcd_step_host_async(self.h, p(actions), p(full_state), p(lcd_bits), p(done)))

  def step_host_wait(self, keep_in_flight=0):
    """AsyncVectorEnv.step_wait: block until at most `keep_in_flight` submitted steps are unfinished"""
    _lib.check(self.l.blcd_step_host_wait(self.h, int(keep_in_flight)))

  def set_bodies(self, bodies, variants=None):
    b = torch.as_tensor(np.ascontiguousarray(bodies, np.float32)).to(self.device)
    v = None if variants is None else torch.as_tensor(np.ascontiguousarray(variants, np.uint32).view(np.int32)).to(self.device)
    assert tuple(b.shape) == (self.n, self.B, 6)
    _lib.check(self.l.blcd_set_bodies(self.h, _ptr(b), _ptr(v), self._stream()))

  def get_bodies(self):
    out = torch.empty((self.n, self.B, 6), dtype=torch.float32, device=self.device)
    _lib.check(self.l.blcd_get_bodies(self.h, _ptr(out), self._stream()))
    return out.cpu().numpy()

  def get_poses_dev(self):
    """(poses [N, B, 4] float32 = x, y, sin, cos; variants [N] int32) exactly as the rasterizer sees them"""
    poses = torch.empty((self.n, self.B, 4), dtype=torch.float32, device=self.device)
    variants = torch.empty((self.n,), dtype=torch.int32, device=self.device)
    _lib.check(self.l.blcd_get_poses(self.h, _ptr(poses), _ptr(variants), self._stream()))
    return poses, variants

  def check_finite(self, auto_reset=False):
    """failure detection: (number of worlds whose state went non-finite, uint8 flag tensor); optionally reset them"""
    flags = torch.zeros((self.n,), dtype=torch.uint8, device=self.device)
    cnt = C.c_int64()
    torch.cuda.current_stream(self.device).synchronize()
    _lib.check(self.l.blcd_check_finite(self.h, _ptr(flags), C.byref(cnt)))
    if auto_reset and cnt.value:
      self.reset_dev(torch.nonzero(flags).flatten().to(torch.int64))
    return int(cnt.value), flags

  def counters(self):
    out = torch.empty((self.n, 8), dtype=torch.int32, device=self.device)
    _lib.check(self.l.blcd_get_counters(self.h, _ptr(out), self._stream()))
    return out.cpu().numpy().view(np.uint32)

  def save_state(self):
    buf = torch.empty((int(self.l.blcd_state_bytes(self.h)),), dtype=torch.uint8, device=self.device)
    _lib.check(self.l.blcd_save_state(self.h, _ptr(buf), self._stream()))
    return buf

  def load_state(self, buf):
    assert buf.numel() == int(self.l.blcd_state_bytes(self.h))
    _lib.check(self.l.blcd_load_state(self.h, _ptr(buf), self._stream()))

  def enable_timing(self, on=True):
    _lib.check(self.l.blcd_enable_timing(self.h, int(on)))

  def last_step_ms(self):
    ms = C.c_float()
    _lib.check(self.l.blcd_last_step_ms(self.h, C.byref(ms)))
    return ms.value

  # -- numpy API with the reference vector-env call shape ----------------------------------------------------------------
  def bits_shape(self, height=None, width=None):
    """trailing shape of a packed frame: [H] words, or [H, 2] for frames wider than 32 px (include/boxlcd_b200.h)"""
    h, w = height or self.H, width or self.W
    return (h,) if w <= 32 else (h, (w + 31) // 32)

  def unpack_lcd(self, bits, width=None):
    """packed rows -> bool [..., H, W] (True = background), the reference's `lcd` observation"""
    w = width or self.W
    # byte k of a little-endian row word holds pixels 8k .. 8k+7: unpack on uint8 views, so the only intermediate is one
    # byte per pixel (an int32 shift-and-mask would need 4-8 bytes per pixel: tens of GB for a dataset-sized batch)
    rows = bits.shape if w <= 32 else bits.shape[:-1]                  # [..., H]
    by = bits.contiguous().view(torch.uint8).reshape(rows + (-1,))     # [..., H, 4 * words]
    shifts = torch.arange(8, device=bits.device, dtype=torch.uint8)
    px = ((by.unsqueeze(-1) >> shifts) & 1).reshape(rows + (-1,))      # [..., H, 32 * words]
    return px[..., :w].to(torch.bool)

  def _obs_numpy(self, obs):
    return {'full_state': obs['full_state'].cpu().numpy(), 'proprio': obs['proprio'].cpu().numpy(),
            'lcd': self.unpack_lcd(obs['lcd_bits']).cpu().numpy()}

  def reset(self, idxs=None, full_state=None, proprio=None):
    """AsyncVectorEnv.reset(idxs, **kwargs): reset the listed worlds (None = all), optionally from states / proprio vectors"""
    self.reset_async(idxs, full_state=full_state, proprio=proprio)
    return self.reset_wait(idxs)

  def observe(self):
    return self._obs_numpy(self.observe_dev())

  def step(self, actions):
    a = torch.as_tensor(np.ascontiguousarray(actions, np.float32).reshape(self.n, self.A)).to(self.device)
    obs, _ = self.step_dev(a, observe=True)
    done = obs['done'].cpu().numpy().astype(bool)
    return self._obs_numpy(obs), np.zeros(self.n), done, [{'timeout': bool(d)} for d in done]

  # -- AsyncVectorEnv's split calls (research/wrappers/async_vector_env.py:131-242) ---------------------------------------
  # *_async enqueues the kernel on the current CUDA stream and returns at once; *_wait copies the result to the host.
  def step_async(self, actions):
    if getattr(self, '_pending', None) is not None:
      raise RuntimeError(f'Calling `step_async` while waiting for a pending call to `{self._pending[0]}` to complete.')
    a = torch.as_tensor(np.ascontiguousarray(actions, np.float32).reshape(self.n, self.A)).to(self.device, non_blocking=True)
    obs, _ = self.step_dev(a, observe=True)
    self._pending = ('step', obs)

  def step_wait(self, timeout=None):
    if getattr(self, '_pending', None) is None or self._pending[0] != 'step':
      raise RuntimeError('Calling `step_wait` without any prior call to `step_async`.')
    obs, self._pending = self._pending[1], None
    done = obs['done'].cpu().numpy().astype(bool)
    return self._obs_numpy(obs), np.zeros(self.n), done, [{'timeout': bool(d)} for d in done]

  def reset_async(self, idxs=None, **kwargs):
    if getattr(self, '_pending', None) is not None:
      raise RuntimeError(f'Calling `reset_async` while waiting for a pending call to `{self._pending[0]}` to complete.')
    full_state, proprio = kwargs.get('full_state'), kwargs.get('proprio')
    if proprio is not None:
      proprio = np.asarray(proprio, np.float32)
      full_state = np.zeros(proprio.shape[:-1] + (self.S,), np.float32)
      full_state[..., self.env.pobs_idxs] = proprio
    idx_t = None if idxs is None else torch.as_tensor(np.asarray(idxs, np.int64)).to(self.device)
    fs_t = None if full_state is None else torch.as_tensor(np.ascontiguousarray(full_state, np.float32)).to(self.device)
    self.reset_dev(idx_t, fs_t)
    self._pending = ('reset', self.observe_dev())

  def reset_wait(self, idxs=None, timeout=None):
    if getattr(self, '_pending', None) is None or self._pending[0] != 'reset':
      raise RuntimeError('Calling `reset_wait` without any prior call to `reset_async`.')
    obs, self._pending = self._obs_numpy(self._pending[1]), None
    if idxs is not None:
      sel = np.asarray(idxs, np.int64)
      obs = {k: v[sel] for k, v in obs.items()}
    return obs

  def render(self, width=None, height=None):
    """lcd_render(width, height) for every world from its current pose -> bool [N, height, width]"""
    width = width or self.W
    height = height or self.H
    if (width, height) == (self.W, self.H):
      return self.unpack_lcd(self.observe_dev()['lcd_bits']).cpu().numpy()
    # the library rejects widths its profile cannot pack (32 px small, 64 px large) with an error message
    poses, variants = self.get_poses_dev()
    bits = self.render_poses_dev(poses, variants, width, height)
    return self.unpack_lcd(bits, width).cpu().numpy()

  def render_rgb(self, width=None, height=None, idxs=None):
    """lcd_render(width, height, lcd_mode='RGB') for the listed worlds (None = all) -> uint8 [n, height, width, 3].
    Host Pillow over the device poses (boxlcd_b200/rgb_render.py)."""
    from boxlcd_b200 import rgb_render
    width = width or self.W
    height = height or self.H
    poses, variants = self.get_poses_dev()
    poses, variants = poses.cpu().numpy(), variants.cpu().numpy()
    sel = range(self.n) if idxs is None else list(idxs)
    cache = {}
    out = np.empty((len(sel), height, width, 3), np.uint8)
    for k, i in enumerate(sel):
      var = int(variants[i])
      if var not in cache:
        cache[var] = rgb_render.body_shapes(self.spec, var)
      out[k] = rgb_render.render_rgb(cache[var], poses[i], self.env.WIDTH, width, height)
    return out


class AsyncVectorEnv(VecWorldEnv):
  """Constructor shape of the reference's `research/wrappers/async_vector_env.AsyncVectorEnv(env_fns, ...)`: callers such as
  `research/data.py:24-29` build `AsyncVectorEnv([env_fn(G) for _ in range(G.num_envs)])`; here the list only tells how many
  worlds to allocate -- one instance is built for the scene description, and all worlds live in one GPU batch (the
  multiprocessing options are accepted and ignored)."""

  def __init__(self, env_fns, observation_space=None, action_space=None, shared_memory=True, copy=True, context=None, daemon=True, worker=None,
               device=None, seed=0):
    env_fns = list(env_fns)
    super().__init__(env_fns[0](), len(env_fns), device=device, seed=seed)
