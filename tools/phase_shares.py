import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import boxlcd_b200 as b
from boxlcd_b200.vec_env import VecWorldEnv
for name,n,T in [('Urchin',75776,30),('UrchinBall',75776,30),('Bounce2',75776,30)]:
    e=b.env_map[name](); v=VecWorldEnv(e,n,seed=0); v.reset_dev(); v.rollout_dev(20)
    c0=v.counters().astype(np.int64); v.rollout_dev(T); torch.cuda.synchronize(); c=(v.counters().astype(np.int64)-c0)
    tot=c[:,:7].sum(1).mean()
    names=['setup(collide,islands,init)','velocity','position','writeback+broadphase','toi','barrier_wait','obs+render']
    print(name, {k: round(100*c[:,i].mean()/tot,1) for i,k in enumerate(names)}, 'cycles/env-step/world(x64)', round(tot/T))
