"""LuxoCube dataset collection end to end -- simulate on the GPU, bring the rollouts to the host in the
reference's array layout, write barrel files.  Reports env-steps/s of each stage, with the parallel writer, without zlib,
and (on a slice) with the reference's single-threaded np.savez_compressed.
  python tools/time_collect.py [ENV] [N_ROLLOUTS]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import boxlcd_b200 as blcd
from boxlcd_b200 import collect
from boxlcd_b200.npz_writer import savez_compressed_parallel
name = sys.argv[1] if len(sys.argv) > 1 else 'LuxoCube'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
env = blcd.env_map[name]()
T = env.G.ep_len
collect.collect_arrays(env, 2000, T, batch=2000)   # warm-up (library load, allocator)
t0 = time.time(); data = collect.collect_arrays(env, n, T, batch=min(n, 65536)); t_sim = time.time() - t0
steps = n * T
nbytes = sum(v.nbytes for v in data.values())
d = tempfile.mkdtemp()
t0 = time.time()
for i in range(0, n, 1000):
  savez_compressed_parallel(os.path.join(d, f'{i}-{T}.barrel'), **{k: v[i:i + 1000] for k, v in data.items()})
t_par = time.time() - t0
size_par = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
t0 = time.time(); np.savez(os.path.join(d, 'raw.npz'), **{k: v[:4000] for k, v in data.items()}); t_raw = (time.time() - t0) * n / 4000
t0 = time.time(); np.savez_compressed(os.path.join(d, 'ref.npz'), **{k: v[:2000] for k, v in data.items()}); t_ref = (time.time() - t0) * n / 2000
print(f'{name}: {n} rollouts x {T} steps = {steps / 1e6:.1f} M env-steps, {nbytes / 1e9:.2f} GB of arrays, {os.cpu_count()} host threads')
print(f'  simulate + device->host + unpack to bool frames : {t_sim:6.2f} s  {steps / t_sim / 1e6:7.2f} M env-steps/s')
print(f'  write {n // 1000} barrels, parallel deflate            : {t_par:6.2f} s  {steps / t_par / 1e6:7.2f} M env-steps/s  ({size_par / 1e6:.0f} MB on disk)')
print(f'  write without zlib (np.savez, extrapolated)      : {t_raw:6.2f} s  {steps / t_raw / 1e6:7.2f} M env-steps/s')
print(f'  reference writer np.savez_compressed (extrapol.) : {t_ref:6.2f} s  {steps / t_ref / 1e6:7.2f} M env-steps/s')
print(f'  end to end with the parallel writer              : {t_sim + t_par:6.2f} s  {steps / (t_sim + t_par) / 1e6:7.2f} M env-steps/s')
