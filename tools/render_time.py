"""frames/s of blcd_render_poses on 4 Mi frames drawn from simulated poses: python tools/render_time.py [env]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import boxlcd_b200 as b
from boxlcd_b200.vec_env import VecWorldEnv
name = sys.argv[1] if len(sys.argv) > 1 else 'Urchin'
v = VecWorldEnv(b.env_map[name](), 65536, seed=0)
v.reset_dev(); v.rollout_dev(20)
poses, variants = v.get_poses_dev()
nr = 4 << 20
idx = torch.randint(0, v.n, (nr,), device='cuda')
poses, variants = poses[idx].contiguous(), variants[idx].contiguous()
for _ in range(3): v.render_poses_dev(poses, variants)
torch.cuda.synchronize()
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): v.render_poses_dev(poses, variants)
c.record(); torch.cuda.synchronize()
ms = a.elapsed_time(c) / 5
print(f'{name}: {nr / ms / 1e6:.3f} G frames/s ({ms:.3f} ms for {nr} frames) rows-kernel={os.environ.get("BLCD_RENDER_ROWS", "0")}', flush=True)
