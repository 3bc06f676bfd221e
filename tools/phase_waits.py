"""barrier wait per phase (diagnostic build -DBLCD_PHASE_CLOCKS=2)"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import boxlcd_b200 as b
from boxlcd_b200.vec_env import VecWorldEnv
for name, n, T in [('Urchin', 75776, 30)]:
  e = b.env_map[name](); v = VecWorldEnv(e, n, seed=0); v.reset_dev(); v.rollout_dev(20)
  c0 = v.counters().astype(np.int64); v.rollout_dev(T); torch.cuda.synchronize(); c = v.counters().astype(np.int64) - c0
  tot = c[:, [0, 1, 2, 3, 4, 6, 7]].sum(1).mean()
  names = ['wait after setup', 'wait after velocity', 'wait after position', 'wait after writeback', 'wait after toi', '-', 'obs+render', 'all work']
  print(name, {k: round(100 * c[:, i].mean() / tot, 1) for i, k in enumerate(names) if k != '-'})
