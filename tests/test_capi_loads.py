"""The C-ABI library must load and export every symbol include/boxlcd_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re
import pytest
from boxlcd_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
  if not os.path.exists(_lib.LIB_PATH):
    _lib.build()
  l = ctypes.CDLL(_lib.LIB_PATH)
  header = open(os.path.join(ROOT, 'include', 'boxlcd_b200.h')).read()
  declared = set(re.findall(r'\b(blcd_[a-z_0-9]+)\s*\(', header))
  assert len(declared) >= 20
  for sym in sorted(declared):
    assert hasattr(l, sym), f'{sym} declared in include/boxlcd_b200.h but not exported'
  assert set(_lib.SYMBOLS) <= declared
  assert l.blcd_version() >= 120


def test_spec_struct_layout_matches_header():
  # sizeof(blcd_spec) as compiled by gcc from the header == ctypes mirror in boxlcd_b200/spec.py
  import subprocess, tempfile
  from boxlcd_b200 import spec
  src = '#include <stdio.h>\n#include "boxlcd_b200.h"\nint main(){printf("%zu %zu %zu %zu", sizeof(blcd_spec), sizeof(blcd_body_def), sizeof(blcd_joint_def), sizeof(blcd_shape_def));return 0;}'
  with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, 't.c'), 'w').write(src)
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), os.path.join(d, 't.c'), '-o', os.path.join(d, 't')], check=True)
    out = subprocess.run([os.path.join(d, 't')], capture_output=True, text=True, check=True).stdout.split()
  assert [int(x) for x in out] == [ctypes.sizeof(spec.Spec), ctypes.sizeof(spec.BodyDef), ctypes.sizeof(spec.JointDef), ctypes.sizeof(spec.ShapeDef)]


def test_no_cpu_fallback_without_gpu():
  import torch
  import boxlcd_b200 as blcd
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  env = blcd.envs.Dropbox()
  with pytest.raises(RuntimeError, match='CUDA'):
    env.reset()
