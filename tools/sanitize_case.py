"""small end-to-end case for compute-sanitizer: reset, fused rollout, step+observe, render, state save/load"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import boxlcd_b200 as b
from boxlcd_b200.vec_env import VecWorldEnv
for name, n in [('Urchin', 300), ('UrchinBall', 130), ('Object2', 70)]:
  env = b.env_map[name]()
  v = VecWorldEnv(env, n, seed=1)
  v.reset_dev()
  v.rollout_dev(3)
  obs, act = v.step_dev(None, observe=True)
  poses, var = v.get_poses_dev()
  v.render_poses_dev(poses, var)
  snap = v.save_state(); v.load_state(snap)
  v.reset_dev(torch.tensor([0, n - 1], device='cuda'))
  torch.cuda.synchronize()
  print(name, 'ok', int(v.counters()[:, 5].sum()))
  v.close()
