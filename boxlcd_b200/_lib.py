"""ctypes binding of libboxlcd_b200.so (include/boxlcd_b200.h).  The library is the product: if it is missing or cannot
be loaded this module raises -- there is no CPU fallback behind it."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BLCD_LIB') or os.path.join(HERE, 'libboxlcd_b200.so')   # BLCD_LIB: experiment builds only
CSRC = os.path.join(HERE, 'csrc')
_lib = None

SYMBOLS = ['blcd_last_error', 'blcd_version', 'blcd_create', 'blcd_destroy', 'blcd_rekey', 'blcd_reset', 'blcd_step', 'blcd_step_observe', 'blcd_observe',
           'blcd_rollout', 'blcd_step_host', 'blcd_pin_host', 'blcd_unpin_host', 'blcd_step_host_async', 'blcd_step_host_wait', 'blcd_render_poses', 'blcd_render_poses_sized', 'blcd_set_bodies', 'blcd_get_bodies', 'blcd_get_poses', 'blcd_check_finite',
           'blcd_state_bytes', 'blcd_save_state', 'blcd_load_state', 'blcd_num_worlds', 'blcd_kernel_launches', 'blcd_last_step_ms',
           'blcd_enable_timing', 'blcd_get_counters', 'blcd_scene_info', 'blcd_measure_peaks']


def build(force=False):
  """compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)"""
  srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h', '.cpp'))] + [os.path.join(HERE, '..', 'include', 'boxlcd_b200.h')]
  stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
  if stale:
    subprocess.run(['make', '-s', '-j3', '-C', CSRC], check=True)
  return LIB_PATH


def lib():
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` (needs nvcc). '
                       'boxlcd_b200 has no CPU fallback.')
  l = C.CDLL(LIB_PATH)
  vp, i64, u64, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32
  l.blcd_last_error.restype = C.c_char_p
  l.blcd_create.argtypes = [vp, i64, C.c_int, u64, i64, C.POINTER(vp)]
  l.blcd_destroy.argtypes = [vp]
  l.blcd_rekey.argtypes = [vp, u64, i64]
  l.blcd_reset.argtypes = [vp, vp, i64, vp, u64]
  l.blcd_step.argtypes = [vp, vp, vp, u64]
  l.blcd_step_observe.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, u64]
  l.blcd_observe.argtypes = [vp, vp, vp, vp, vp, vp, u64]
  l.blcd_rollout.argtypes = [vp, i32, vp, vp, vp, u64]
  l.blcd_step_host.argtypes = [vp, vp, vp, vp, vp]
  l.blcd_pin_host.argtypes = [vp, vp, i64]
  l.blcd_unpin_host.argtypes = [vp, vp]
  l.blcd_step_host_async.argtypes = [vp, vp, vp, vp, vp]
  l.blcd_step_host_wait.argtypes = [vp, i32]
  l.blcd_render_poses.argtypes = [vp, vp, vp, i64, vp, u64]
  l.blcd_render_poses_sized.argtypes = [vp, vp, vp, i64, i32, i32, vp, u64]
  l.blcd_set_bodies.argtypes = [vp, vp, vp, u64]
  l.blcd_get_bodies.argtypes = [vp, vp, u64]
  l.blcd_get_poses.argtypes = [vp, vp, vp, u64]
  l.blcd_check_finite.argtypes = [vp, vp, C.POINTER(i64)]
  l.blcd_state_bytes.argtypes = [vp]
  l.blcd_state_bytes.restype = i64
  l.blcd_save_state.argtypes = [vp, vp, u64]
  l.blcd_load_state.argtypes = [vp, vp, u64]
  l.blcd_num_worlds.argtypes = [vp]
  l.blcd_num_worlds.restype = i64
  l.blcd_kernel_launches.argtypes = [vp]
  l.blcd_kernel_launches.restype = i64
  l.blcd_last_step_ms.argtypes = [vp, C.POINTER(C.c_float)]
  l.blcd_enable_timing.argtypes = [vp, C.c_int]
  l.blcd_get_counters.argtypes = [vp, vp, u64]
  l.blcd_scene_info.argtypes = [vp, vp]
  l.blcd_measure_peaks.argtypes = [C.c_int, C.POINTER(C.c_double)]
  _lib = l
  return l


def check(rc):
  if rc != 0:
    raise RuntimeError('libboxlcd_b200: ' + lib().blcd_last_error().decode())
