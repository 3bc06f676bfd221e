"""GPU probe: which of the two device paths (fused kernel / phase pipeline) agrees with the CPU oracle, per counter and per step"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import boxlcd_b200 as blcd
from boxlcd_b200.vec_env import VecWorldEnv
from oracle import oracle
name = sys.argv[1] if len(sys.argv) > 1 else 'Urchin'
m, T = 4096, 6
env = blcd.env_map[name]()
ow = oracle.OracleWorlds(env.layout.spec, m, seed=0, threads=16); ow.reset()
ro = ow.rollout(T)
co = ow.counters()
for pipeline in (0, 1):
  os.environ['BLCD_PIPELINE'] = str(pipeline)
  v = VecWorldEnv(env, m, seed=0)
  v.reset_dev()
  r = v.rollout_dev(T)
  fs = r['full_state'].cpu().numpy()
  err = np.abs(fs - ro['full_state']).max(2)    # [m, T]
  c = v.counters()
  print(f'{name} pipeline={pipeline}: worlds within 1e-5 of the oracle at t=1..{T - 1}:', [round(float((err[:, t] < 1e-5).mean()), 4) for t in range(1, T)],
        'counters differing from the oracle per column:', (c != co).sum(0).tolist(), flush=True)
