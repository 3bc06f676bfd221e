"""AddressSanitizer + UBSan over the CUDA path's simulation source (host build): every env (both scene-size profiles), 30-step rollouts plus a
reset-from-state and a step.  Stands in for compute-sanitizer, which is closed on the GPU pool."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_simulation_source_is_clean_under_asan_ubsan():
  subprocess.run(['make', '-s', '-C', os.path.join(ROOT, 'tests', 'hostsim'), '_build/libhostsim_asan.so', '_build/libhostsim_large_asan.so'], check=True)
  asan = subprocess.run(['g++', '-print-file-name=libasan.so'], capture_output=True, text=True).stdout.strip()
  ubsan = subprocess.run(['g++', '-print-file-name=libubsan.so'], capture_output=True, text=True).stdout.strip()
  env = dict(os.environ, LD_PRELOAD=f'{asan} {ubsan}', ASAN_OPTIONS='detect_leaks=0', UBSAN_OPTIONS='halt_on_error=1')
  r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'asan_hostsim_case.py')], env=env, capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stderr[-3000:]
  assert 'runtime error' not in r.stderr and 'AddressSanitizer' not in r.stderr, r.stderr[-3000:]
  assert r.stdout.count(' ok True 0') == 18, r.stdout
