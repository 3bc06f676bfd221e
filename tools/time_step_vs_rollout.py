"""Device-resident comparison of 100 x k_step (one launch per env step) with one k_rollout launch of 100 steps."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, boxlcd_b200 as blcd
from boxlcd_b200.vec_env import VecWorldEnv
name, n = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ('Urchin', 262144)
env = blcd.env_map[name]()
v = VecWorldEnv(env, n, seed=0)
T = 100
def ev(fn):
  torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record(); fn(); b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)
for rep in range(2):
  v.reset_dev(); t_roll = ev(lambda: v.rollout_dev(T))
  v.reset_dev(); t_step = ev(lambda: [v.step_dev(None, observe=True) for _ in range(T)])
  v.reset_dev(); t_noobs = ev(lambda: [v.step_dev(None, observe=False) for _ in range(T)])
  print(f'{name} n={n}: rollout {t_roll / T:.3f} ms/step, step+observe launches {t_step / T:.3f} ms/step, step only {t_noobs / T:.3f} ms/step', flush=True)
