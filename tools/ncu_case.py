"""The launch captured by `ncu --set full` for profiles/r01*_rollout_ncu_raw.csv: Urchin, 37 888 worlds (148 SMs x 256), a 20-step
warm-up rollout (robots settled on the floor), then ONE 3-step k_rollout launch (the second k_rollout of the process)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import boxlcd_b200 as b
from boxlcd_b200.vec_env import VecWorldEnv
name = sys.argv[1] if len(sys.argv) > 1 else 'Urchin'
v = VecWorldEnv(b.env_map[name](), 37888, seed=0)
v.reset_dev()
v.rollout_dev(20)
v.rollout_dev(3)
torch.cuda.synchronize()
print('ok', int(v.counters()[:, 5].sum()))
