"""per-warp distribution of phase work inside a block (diagnostic build -DBLCD_PHASE_CLOCKS)"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import boxlcd_b200 as b
from boxlcd_b200.vec_env import VecWorldEnv
e = b.envs.Urchin(); n = 75776; v = VecWorldEnv(e, n, seed=0); v.reset_dev(); v.rollout_dev(20)
names = ['setup', 'velocity', 'position', 'writeback', 'toi']
c0 = v.counters().astype(np.int64); v.rollout_dev(1); torch.cuda.synchronize(); c = v.counters().astype(np.int64) - c0
for i, k in enumerate(names):
  w = c[::32, i].reshape(-1, 8).astype(np.float64)     # [blocks, warps]: cycles (x64) of one env step = 3 sub-steps
  print(f'{k:10s} mean {w.mean():9.0f}  block-max/mean {np.mean(w.max(1)) / w.mean():.2f}  p99/mean {np.percentile(w, 99) / w.mean():.2f}  cv {w.std() / w.mean():.2f}')
oc = v.counters().astype(np.int64)
