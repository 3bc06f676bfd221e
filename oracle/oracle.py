"""ctypes front-end of oracle/_build/liboracle.so (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

  lcd_render(...)      -> oracle/lcd_oracle.c   restatement of WorldEnv.lcd_render (world_env.py:460-512)
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_build', 'liboracle.so')
_lib = None


class LcdShape(C.Structure):
  _fields_ = [('kind', C.c_int32), ('n', C.c_int32), ('radius', C.c_float), ('verts', (C.c_float * 2) * 8)]


LCD_SHAPE_DTYPE = np.dtype([('kind', np.int32), ('n', np.int32), ('radius', np.float32), ('verts', np.float32, (8, 2))])
assert LCD_SHAPE_DTYPE.itemsize == C.sizeof(LcdShape)


def build(force=False):
  """(re)compile the oracle with make; sources newer than the .so trigger a rebuild"""
  srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.c', '.cpp', '.h'))] + [os.path.join(HERE, '..', 'include', 'boxlcd_b200.h')]
  stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs if os.path.exists(s))
  if stale:
    subprocess.run(['make', '-s', '-C', HERE], check=True)
  return LIB_PATH


def lib():
  global _lib
  if _lib is None:
    build()
    _lib = C.CDLL(LIB_PATH)
    _lib.blcd_oracle_lcd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    _lib.blcd_oracle_lcd.restype = None
  return _lib


def make_shapes(kind, nvert, radius, verts):
  """arrays [..., B] / [..., B, 8, 2] -> structured lcd_shape array of the same leading shape"""
  kind = np.asarray(kind)
  out = np.zeros(kind.shape, LCD_SHAPE_DTYPE)
  out['kind'], out['n'], out['radius'], out['verts'] = kind, nvert, radius, verts
  return out


def lcd_render(shapes, poses, world_w, lcd_w, lcd_h, rules=0):
  """shapes: structured array [B] (shared) or [n, B] (per world); poses [n, B, 4] f32 (x, y, sin, cos).
  Returns bits [n, lcd_h] uint32 (bit x = pixel x, 1 = background, row 0 = top)."""
  poses = np.ascontiguousarray(poses, np.float32)
  n, B = poses.shape[0], poses.shape[1]
  shapes = np.ascontiguousarray(shapes)
  per_world = int(shapes.ndim == 2)
  assert shapes.shape[-1] == B and (not per_world or shapes.shape[0] == n)
  bits = np.zeros((n,) + bits_shape(lcd_h, lcd_w), np.uint32)
  lib().blcd_oracle_lcd(shapes.ctypes.data, per_world, B, poses.ctypes.data, n, world_w, lcd_w, lcd_h, rules, bits.ctypes.data)
  return bits


def bits_shape(lcd_h, lcd_w):
  """trailing shape of a bit-packed frame: [H] uint32, or [H, 2] for frames wider than 32 px (include/boxlcd_b200.h)"""
  return (lcd_h,) if lcd_w <= 32 else (lcd_h, (lcd_w + 31) // 32)


def unpack_bits(bits, lcd_w):
  """[..., H] (or [..., H, 2]) uint32 -> bool [..., H, W] (the reference's lcd array; True = background)"""
  bits = np.asarray(bits)
  px = ((bits[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
  if lcd_w > 32:
    px = px.reshape(px.shape[:-2] + (-1,))
  return px[..., :lcd_w]


# ----------------------------------------------------------------------------------------------------------------------
# b2_oracle.cpp: batches of independent worlds (restatement of WorldEnv.reset/step/_get_obs + b2World.Step)
BODY_STATE = 6
N_COUNTERS = 8
COUNTER_NAMES = ['contacts', 'pos_iters', 'toi_events', 'toi_calls', 'sleep_steps', 'overflow', 'manifold_points', 'substeps']


def _p(a):
  return None if a is None else a.ctypes.data_as(C.c_void_p)


def _bind_worlds(l):
  if getattr(l, '_worlds_bound', False):
    return l
  vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
  l.blcd_oracle_worlds_new.argtypes = [vp, i64, C.c_uint64, i64]
  l.blcd_oracle_worlds_new.restype = vp
  l.blcd_oracle_worlds_free.argtypes = [vp]
  l.blcd_oracle_worlds_select.argtypes = [vp, vp, i64]
  l.blcd_oracle_worlds_select.restype = vp
  l.blcd_oracle_worlds_reset.argtypes = [vp, vp, i64, vp, i32]
  l.blcd_oracle_worlds_set_bodies.argtypes = [vp, vp, vp, i32]
  l.blcd_oracle_worlds_get_bodies.argtypes = [vp, vp]
  l.blcd_oracle_worlds_get_poses.argtypes = [vp, vp, vp]
  l.blcd_oracle_worlds_step.argtypes = [vp, vp, vp, i32]
  l.blcd_oracle_worlds_observe.argtypes = [vp, vp, vp, vp, vp, i32]
  l.blcd_oracle_worlds_rollout.argtypes = [vp, i32, vp, vp, vp, i32]
  l.blcd_oracle_worlds_counters.argtypes = [vp, vp]
  l.blcd_oracle_worlds_lcd_shapes.argtypes = [vp, i64, vp]
  l.blcd_oracle_worlds_mass.argtypes = [vp, i64, i32, vp]
  l._worlds_bound = True
  return l


class OracleWorlds:
  """n independent CPU worlds of one scene.  `spec` is a boxlcd_b200.spec.Spec (the same POD the CUDA library takes)."""

  def __init__(self, spec, n, seed=0, world_offset=0, threads=1):
    self.l = _bind_worlds(lib())
    self.spec, self.n, self.threads = spec, int(n), int(threads)
    self.h = self.l.blcd_oracle_worlds_new(C.byref(spec), self.n, seed, world_offset)
    self.B, self.S, self.A = spec.n_bodies, spec.obs_size, spec.act_size
    self.P = max(spec.pobs_size, 1)
    self.H, self.W = spec.lcd_h, spec.lcd_w

  def close(self):
    if self.h:
      self.l.blcd_oracle_worlds_free(self.h)
      self.h = None

  __del__ = close

  def select(self, idx):
    """new OracleWorlds holding copies (full simulation state) of the listed worlds"""
    idx = np.ascontiguousarray(idx, np.int64)
    new = object.__new__(OracleWorlds)
    new.__dict__.update(self.__dict__)
    new.n = len(idx)
    new.h = self.l.blcd_oracle_worlds_select(self.h, _p(idx), len(idx))
    return new

  def reset(self, idx=None, full_state=None):
    idx_a = None if idx is None else np.ascontiguousarray(idx, np.int64)
    fs = None if full_state is None else np.ascontiguousarray(full_state, np.float32)
    self.l.blcd_oracle_worlds_reset(self.h, _p(idx_a), 0 if idx_a is None else len(idx_a), _p(fs), self.threads)

  def set_bodies(self, bodies, variants=None):
    bodies = np.ascontiguousarray(bodies, np.float32)
    assert bodies.shape == (self.n, self.B, BODY_STATE)
    v = None if variants is None else np.ascontiguousarray(variants, np.uint32)
    self.l.blcd_oracle_worlds_set_bodies(self.h, _p(bodies), _p(v), self.threads)

  def get_bodies(self):
    out = np.zeros((self.n, self.B, BODY_STATE), np.float32)
    self.l.blcd_oracle_worlds_get_bodies(self.h, _p(out))
    return out

  def get_poses(self):
    poses = np.zeros((self.n, self.B, 4), np.float32)
    variants = np.zeros(self.n, np.uint32)
    self.l.blcd_oracle_worlds_get_poses(self.h, _p(poses), _p(variants))
    return poses, variants

  def step(self, actions=None):
    a = None if actions is None else np.ascontiguousarray(actions, np.float32)
    out = np.zeros((self.n, self.A), np.float32)
    self.l.blcd_oracle_worlds_step(self.h, _p(a), _p(out), self.threads)
    return out

  def observe(self):
    fs = np.zeros((self.n, self.S), np.float32)
    pr = np.zeros((self.n, self.P), np.float32)
    bits = np.zeros((self.n,) + bits_shape(self.H, self.W), np.uint32)
    done = np.zeros(self.n, np.uint8)
    self.l.blcd_oracle_worlds_observe(self.h, _p(fs), _p(pr), _p(bits), _p(done), self.threads)
    return {'full_state': fs, 'proprio': pr, 'lcd_bits': bits, 'done': done.astype(bool)}

  def rollout(self, T, want=('full_state', 'lcd_bits', 'action')):
    fs = np.zeros((self.n, T, self.S), np.float32) if 'full_state' in want else None
    bits = np.zeros((self.n, T) + bits_shape(self.H, self.W), np.uint32) if 'lcd_bits' in want else None
    act = np.zeros((self.n, T, self.A), np.float32) if 'action' in want else None
    self.l.blcd_oracle_worlds_rollout(self.h, T, _p(fs), _p(bits), _p(act), self.threads)
    return {'full_state': fs, 'lcd_bits': bits, 'action': act}

  def counters(self):
    out = np.zeros((self.n, N_COUNTERS), np.uint32)
    self.l.blcd_oracle_worlds_counters(self.h, _p(out))
    return out

  def lcd_shapes(self, i=0):
    out = np.zeros(self.B, LCD_SHAPE_DTYPE)
    self.l.blcd_oracle_worlds_lcd_shapes(self.h, i, _p(out))
    return out

  def mass(self, k, i=0):
    out = np.zeros(4, np.float32)
    self.l.blcd_oracle_worlds_mass(self.h, i, k, _p(out))
    return out


def place_children(spec, pose_in):
  """world_env.py:230-252 on given root / object poses [B, 3] (float64) -> all poses [B, 3] float32"""
  l = _bind_worlds(lib())
  l.blcd_oracle_place_children.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
  pin = np.ascontiguousarray(pose_in, np.float64)
  out = np.zeros((spec.n_bodies, 3), np.float32)
  l.blcd_oracle_place_children(C.byref(spec), _p(pin), _p(out))
  return out
