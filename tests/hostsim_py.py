"""ctypes wrapper of tests/hostsim (the CUDA path's simulation source compiled for the host; TEST TOOL ONLY)."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hostsim')
LIB = os.path.join(HERE, '_build', 'libhostsim.so')               # small profile (boxlcd_b200/csrc/blcd_profile.h)
LIB_LARGE = os.path.join(HERE, '_build', 'libhostsim_large.so')   # large profile
_libs = {}


def fits_small(spec):
  """same rule as blcd_create (boxlcd_b200/csrc/blcd_dispatch.cpp)"""
  return spec.n_bodies <= 8 and spec.n_joints <= 7 and spec.lcd_w <= 32


def lib(large=False):
  path = LIB_LARGE if large else LIB
  if path not in _libs:
    if path.endswith(('libhostsim.so', 'libhostsim_large.so')):
      subprocess.run(['make', '-s', '-C', HERE], check=True)
    l = C.CDLL(path)
    vp, i64 = C.c_void_p, C.c_int64
    l.hostsim_new.argtypes = [vp, i64, C.c_uint64, i64, C.c_int]
    l.hostsim_new.restype = vp
    l.hostsim_free.argtypes = [vp]
    l.hostsim_reset.argtypes = [vp, vp]
    l.hostsim_set_bodies.argtypes = [vp, vp, vp]
    l.hostsim_get_bodies.argtypes = [vp, vp]
    l.hostsim_step.argtypes = [vp, vp, vp]
    l.hostsim_observe.argtypes = [vp, vp, vp]
    l.hostsim_rollout.argtypes = [vp, C.c_int, vp, vp, vp]
    l.hostsim_step_pipeline.argtypes = [vp, vp, vp]
    l.hostsim_polygon_row_check.argtypes = [i64, C.c_uint64]
    l.hostsim_polygon_row_check.restype = i64
    l.hostsim_rollout_pipeline.argtypes = [vp, C.c_int, vp, vp, vp]
    l.hostsim_counters.argtypes = [vp, vp]
    l.hostsim_render_poses.argtypes = [vp, vp, vp, i64, C.c_int, C.c_int, vp]
    _libs[path] = l
  return _libs[path]


def _p(a):
  return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostSim:
  def __init__(self, spec, n, seed=0, world_offset=0, maxm=0, profile=None):
    self.l = lib(large=(profile == 'large') or (profile is None and not fits_small(spec)))
    self.spec, self.n = spec, int(n)
    self.h = self.l.hostsim_new(C.byref(spec), self.n, seed, world_offset, maxm)
    assert self.h, 'scene rejected'
    self.B, self.S, self.A, self.H, self.W = spec.n_bodies, spec.obs_size, spec.act_size, spec.lcd_h, spec.lcd_w

  def _rows(self, h=None, w=None):
    h, w = h or self.H, w or self.W
    return (h,) if w <= 32 else (h, (w + 31) // 32)

  def __del__(self):
    if getattr(self, 'h', None):
      self.l.hostsim_free(self.h)
      self.h = None

  def reset(self, full_state=None):
    fs = None if full_state is None else np.ascontiguousarray(full_state, np.float32)
    self.l.hostsim_reset(self.h, _p(fs))

  def set_bodies(self, bodies, variants=None):
    b = np.ascontiguousarray(bodies, np.float32)
    v = None if variants is None else np.ascontiguousarray(variants, np.uint32)
    self.l.hostsim_set_bodies(self.h, _p(b), _p(v))

  def get_bodies(self):
    out = np.zeros((self.n, self.B, 6), np.float32)
    self.l.hostsim_get_bodies(self.h, _p(out))
    return out

  def step(self, actions=None, pipeline=False):
    """pipeline=True: the phase pipeline (csrc/blcd_pipeline.cuh), every phase in a fresh poisoned Sim, records through scratch"""
    a = None if actions is None else np.ascontiguousarray(actions, np.float32)
    out = np.zeros((self.n, self.A), np.float32)
    (self.l.hostsim_step_pipeline if pipeline else self.l.hostsim_step)(self.h, _p(a), _p(out))
    return out

  def observe(self):
    fs = np.zeros((self.n, self.S), np.float32)
    bits = np.zeros((self.n,) + self._rows(), np.uint32)
    self.l.hostsim_observe(self.h, _p(fs), _p(bits))
    return {'full_state': fs, 'lcd_bits': bits}

  def rollout(self, T, pipeline=False):
    fs = np.zeros((self.n, T, self.S), np.float32)
    bits = np.zeros((self.n, T) + self._rows(), np.uint32)
    act = np.zeros((self.n, T, self.A), np.float32)
    (self.l.hostsim_rollout_pipeline if pipeline else self.l.hostsim_rollout)(self.h, T, _p(fs), _p(bits), _p(act))
    return {'full_state': fs, 'lcd_bits': bits, 'action': act}

  def counters(self):
    out = np.zeros((self.n, 8), np.uint32)
    self.l.hostsim_counters(self.h, _p(out))
    return out

  def render_poses(self, poses, variants=None, lcd_w=0, lcd_h=0):
    poses = np.ascontiguousarray(poses, np.float32)
    n = poses.shape[0]
    v = None if variants is None else np.ascontiguousarray(variants, np.uint32)
    bits = np.zeros((n,) + self._rows(lcd_h or self.H, lcd_w or self.W), np.uint32)
    self.l.hostsim_render_poses(self.h, _p(poses), _p(v), n, lcd_w, lcd_h, _p(bits))
    return bits
