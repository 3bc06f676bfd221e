import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, boxlcd_b200 as blcd
from boxlcd_b200.vec_env import VecWorldEnv
env = blcd.env_map['Urchin']()
n = 262144
v = VecWorldEnv(env, n, seed=0)
act = np.random.RandomState(0).uniform(-1, 1, (n, 3)).astype(np.float32)
fs = np.zeros((n, v.S), np.float32); bits = np.zeros((n, v.H), np.uint32); dn = np.zeros(n, np.uint8)
v.pin_host(act, fs, bits, dn)
def run(label, fn, T=50):
  v.reset_dev(); [fn() for _ in range(3)]
  torch.cuda.synchronize(); t0 = time.perf_counter()
  [fn() for _ in range(T)]
  torch.cuda.synchronize(); print(label, os.environ.get('BLCD_HOST_CHUNKS'), round((time.perf_counter() - t0) / T * 1e3, 3), 'ms/step', flush=True)
run('no copies      ', lambda: v.step_host(None))
run('actions only   ', lambda: v.step_host(act))
run('acts+fs        ', lambda: v.step_host(act, fs))
run('acts+fs+bits+dn', lambda: v.step_host(act, fs, bits, dn))
# raw copy bandwidth
d = torch.empty(33_554_432 // 4, dtype=torch.float32, device='cuda'); h = torch.empty(33_554_432 // 4, dtype=torch.float32).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print('D2H pinned GB/s', 20 * 33.554432e-3 / (time.perf_counter() - t0))
hb = torch.from_numpy(bits.view(np.int32))
t0 = time.perf_counter()
for _ in range(20): hb.copy_(d[:hb.numel()].view(torch.int32).view(hb.shape), non_blocking=True)
torch.cuda.synchronize(); print('D2H registered numpy GB/s', 20 * hb.numel() * 4e-9 / (time.perf_counter() - t0))
