"""The launch(es) profiled by ncu: `python tools/ncu_case.py [env] [worlds] [T] [range]`.

Default (profiles/r01*_rollout_ncu_raw.csv): Urchin, 37 888 worlds (148 SMs x 256), a 20-step warm-up rollout (robots settled
on the floor), then ONE 3-step rollout.  With a fourth argument `range` the measured rollout is bracketed by
cudaProfilerStart / cudaProfilerStop, so `ncu --profile-from-start off` sees exactly the kernels of that rollout call
(bench.py's instruction-count leg sums their counters)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import boxlcd_b200 as b
from boxlcd_b200.vec_env import VecWorldEnv
name = sys.argv[1] if len(sys.argv) > 1 else 'Urchin'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 37888
T = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ranged = len(sys.argv) > 4 and sys.argv[4] == 'range'
v = VecWorldEnv(b.env_map[name](), n, seed=0)
v.reset_dev()
v.rollout_dev(20)
torch.cuda.synchronize()
if ranged:
  torch.cuda.cudart().cudaProfilerStart()
v.rollout_dev(T)
torch.cuda.synchronize()
if ranged:
  torch.cuda.cudart().cudaProfilerStop()
print('ok', int(v.counters()[:, 5].sum()))
