"""Time one device-resident rollout: python tools/time_rollout.py ENV WORLDS (honours BLCD_BLOCK / BLCD_ALIGN / BLCD_PROFILE)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import os, sys, torch, boxlcd_b200 as blcd
from boxlcd_b200.vec_env import VecWorldEnv
name, n = sys.argv[1], int(sys.argv[2])
env = blcd.env_map[name]()
v = VecWorldEnv(env, n, seed=0)
v.enable_timing(True)
v.reset_dev(); v.rollout_dev(20)
v.reset_dev(); v.rollout_dev(100); torch.cuda.synchronize()
ms = v.last_step_ms()
print(name, 'block', v.info()['block'], 'align', os.environ.get('BLCD_ALIGN'), 'worlds', n, 'M env-steps/s', round(n*100/ms*1e3/1e6, 3), flush=True)
