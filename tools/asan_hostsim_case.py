"""Runs the CUDA path's simulation source, compiled for the host with -fsanitize=address,undefined, over rollouts of every
env (compute-sanitizer is closed on the GPU pool; this covers the same source for out-of-bounds and UB).  Launched by
tests/test_hostsim_asan.py with libasan preloaded."""
import sys, ctypes as C, numpy as np
import os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import hostsim_py
hostsim_py.LIB = os.path.join(ROOT, 'tests', 'hostsim', '_build', 'libhostsim_asan.so')
hostsim_py.LIB_LARGE = os.path.join(ROOT, 'tests', 'hostsim', '_build', 'libhostsim_large_asan.so')
import boxlcd_b200 as b
from hostsim_py import HostSim
assert hostsim_py.lib().hostsim_polygon_row_check(200000, 3) == 0     # the render kernel's unrolled scanline rules
for name in sorted(b.env_map):
    e=b.env_map[name]()
    hs=HostSim(e.layout.spec, 24, seed=2); hs.reset(); r=hs.rollout(30)
    fs=r['full_state'][:, -1]
    hs2=HostSim(e.layout.spec, 24, seed=9); hs2.reset(full_state=fs); hs2.step()
    hp=HostSim(e.layout.spec, 8, seed=2); hp.reset(); rp=hp.rollout(12, pipeline=True)    # the phase pipeline's spill / fill paths (blcd_pipeline.cuh)
    assert (rp['full_state'] == r['full_state'][:8, :12]).all()
    print(name, 'ok', np.isfinite(r['full_state']).all(), int(hs.counters()[:,5].sum()))
