"""the launch profiled for the rasterizer: 1 M Urchin frames from simulated poses through blcd_render_poses"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import boxlcd_b200 as b
from boxlcd_b200.vec_env import VecWorldEnv
name = sys.argv[1] if len(sys.argv) > 1 else 'Urchin'
v = VecWorldEnv(b.env_map[name](), 65536, seed=0)
v.reset_dev(); v.rollout_dev(20)
poses, variants = v.get_poses_dev()
poses = poses[torch.randint(0, v.n, (1 << 20,), device='cuda')].contiguous()
v.render_poses_dev(poses)
torch.cuda.synchronize()
