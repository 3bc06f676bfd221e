"""time the pipeline's rollout only: python tools/pipe_time.py [env] [worlds] [T]  (BLCD_* knobs from the environment)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import boxlcd_b200 as blcd
from boxlcd_b200.vec_env import VecWorldEnv
name = sys.argv[1] if len(sys.argv) > 1 else 'Urchin'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
os.environ.setdefault('BLCD_PIPELINE', '1')
v = VecWorldEnv(blcd.env_map[name](), n, seed=0)
v.reset_dev(); v.rollout_dev(3)
best = 1e9
for rep in range(3):
  v.reset_dev()
  torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record(); v.rollout_dev(T); b.record(); torch.cuda.synchronize()
  best = min(best, a.elapsed_time(b))
v.reset_dev()
r = v.rollout_dev(min(T, 8))
import hashlib
digest = hashlib.sha1(r['full_state'].cpu().numpy().tobytes() + r['lcd_bits'].cpu().numpy().tobytes() + v.counters().tobytes()).hexdigest()[:12]
knobs = {k: os.environ[k] for k in os.environ if k.startswith('BLCD_')}
knobs['sha1'] = digest
print(f'{name} n={n} T={T} {knobs}: {best / T:.3f} ms/env-step, {n * T / best / 1e3:.2f} M env-steps/s', flush=True)
