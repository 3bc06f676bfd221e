"""bench.py -- env-steps/sec (each step includes one LCD frame, full_state and the action) for the B200-native boxLCD
hot path, next to the CPU path on the box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--env Urchin] [--worlds 262144] [--T 100] [--impl reference]

A "step" is one pass of the hot path over one batch: every world of this rank's shard runs a T-env-step random-action
rollout (examples/collect.py:31-39) in ONE launch of the fused kernel (k_rollout: action RNG -> WorldEnv.step -> _get_obs
-> lcd_render, outputs written straight to HBM).  value = env-steps of all ranks / max-over-ranks device time.
e2e = the same count through blcd_step_host (host action buffer in, host observation buffers out, every env step).
Under torchrun each rank owns `--worlds` worlds (weak scaling, no collective: worlds are independent).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES = lambda sp: 4 * sp.lcd_h * ((sp.lcd_w + 31) // 32) + 4 * sp.obs_size + 4 * sp.act_size   # packed frame + full_state + action, SURVEY 8(d)


def parse():
  p = argparse.ArgumentParser()
  p.add_argument('--gpus', type=int, default=1)
  p.add_argument('--steps', type=int, default=3)
  p.add_argument('--warmup', type=int, default=3)
  p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  p.add_argument('--env', default='Urchin')
  p.add_argument('--worlds', type=int, default=262144, help='worlds per GPU')
  p.add_argument('--T', type=int, default=0, help='env steps per rollout (0 = the env\'s ep_len)')
  p.add_argument('--e2e_worlds', type=int, default=0, help='worlds for the host-buffer e2e leg (0 = same as --worlds)')
  p.add_argument('--cpu_seconds', type=float, default=12.0, help='target CPU work for the cpu_baseline sample')
  p.add_argument('--no_cpu', action='store_true')
  p.add_argument('--no_e2e', action='store_true')
  return p.parse_args()


class ClockSampler:
  """samples nvidia-smi clocks / throttle reasons of one GPU every 200 ms while the timed region runs"""
  Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

  def __init__(self, index):
    self.index, self.rows, self.proc = index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200', '-i', str(self.index)],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.th = threading.Thread(target=self._read, daemon=True)
      self.th.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([x.strip() for x in line.split(',')])

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    for r in self.rows:
      try:
        sm.append(float(r[0])); mx.append(float(r[1]))
      except Exception:
        continue
      for name, val in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[3:7]):
        if val.lower().startswith('active'):
          reasons.add(name)
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons), 'samples': len(sm)}


def host_cores():
  try:
    return len(os.sched_getaffinity(0))
  except Exception:
    return os.cpu_count() or 1


def cpu_rollout_rate(spec, T, cores, seconds):
  """times the CPU oracle (oracle/b2_oracle.cpp: the restatement of pybox2d + PIL this repo checks against, kind
  "port") on `cores` threads over a bounded sample of the same workload: reset + T-step random-action rollouts with
  frames.  Returns (env-steps/s, sample description)."""
  from oracle import oracle
  probe_n = cores * 4
  ow = oracle.OracleWorlds(spec, probe_n, seed=0, threads=cores)
  ow.reset()
  t0 = time.perf_counter()
  ow.rollout(min(T, 20), want=('full_state', 'lcd_bits', 'action'))
  rate = probe_n * min(T, 20) / (time.perf_counter() - t0)
  n = int(max(cores, min(65536, rate * seconds / T)))
  n = (n // cores) * cores
  ow = oracle.OracleWorlds(spec, n, seed=1, threads=cores)
  t0 = time.perf_counter()
  ow.reset()
  ow.rollout(T, want=('full_state', 'lcd_bits', 'action'))
  dt = time.perf_counter() - t0
  return n * T / dt, f'{n} worlds x {T} env-steps (reset + rollout + frames), {dt:.1f} s on {cores} threads'


def run_reference(a, rank, world_size):
  """--impl reference: the reference's CPU path for the same workload.  pybox2d is not installable in this image, so
  the arm runs the repo's CPU restatement of it (oracle/, kind "port") on all host threads -- never presented as pybox2d."""
  if rank != 0:
    return
  import boxlcd_b200 as blcd
  env = blcd.env_map[a.env]()
  sp = env.layout.spec
  T = a.T or env.G.ep_len
  cores = host_cores()
  from oracle import oracle
  per_step = max(cores, int(cores * 9000 * 6.0 / T) // cores * cores)   # ~6 s of CPU work per step on ~9k steps/s/core
  times = []
  for it in range(a.warmup + a.steps):
    ow = oracle.OracleWorlds(sp, per_step, seed=it, threads=cores)
    t0 = time.perf_counter()
    ow.reset()
    ow.rollout(T, want=('full_state', 'lcd_bits', 'action'))
    dt = time.perf_counter() - t0
    if it >= a.warmup:
      times.append(dt)
    if it == 0 and dt > 20.0:    # slow box: shrink the sample so the whole run ends within minutes
      per_step = max(cores, int(per_step * 8.0 / dt) // cores * cores)
  ms = 1e3 * float(np.mean(times))
  value = per_step * T / (ms / 1e3)
  line = {
      'impl': 'reference', 'metric': 'env-steps/sec incl. LCD frames', 'value': value, 'unit': 'env-steps/s', 'n_gpus': a.gpus, 'steps': a.steps,
      'warmup': a.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': f'envs.{a.env}() {sp.lcd_h}x{sp.lcd_w}, {a.worlds} worlds per GPU, random-action {T}-step rollouts (reset + step + obs + frame)',
                 'worlds_per_gpu': a.worlds, 'sample_worlds_per_step': per_step, 'T': T, 'note': 'CPU restatement of Box2D 2.3 + PIL rasterizer (oracle/), not pybox2d: pybox2d is not installable here'},
      'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': f'{per_step} worlds x {T} env-steps per step'},
      'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(line), flush=True)


def main():
  a = parse()
  rank = int(os.environ.get('RANK', 0))
  world_size = int(os.environ.get('WORLD_SIZE', 1))
  local_rank = int(os.environ.get('LOCAL_RANK', 0))
  if a.impl == 'reference':
    run_reference(a, rank, world_size)
    return
  import torch
  import torch.distributed as dist
  import boxlcd_b200 as blcd
  from boxlcd_b200.vec_env import VecWorldEnv
  assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback in the product path)'
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world_size > 1:
    dist.init_process_group('nccl', device_id=dev)
  env = blcd.env_map[a.env]()
  sp = env.layout.spec
  T = a.T or env.G.ep_len
  n = a.worlds
  v = VecWorldEnv(env, n, device=dev, seed=0, world_offset=rank * n)
  info = v.info()
  out = v.rollout_dev(1)  # allocate nothing big yet; touch the path once
  f32 = dict(dtype=torch.float32, device=dev)
  fs = torch.empty((n, T, v.S), **f32)
  bits = torch.empty((n, T) + v.bits_shape(), dtype=torch.int32, device=dev)
  act = torch.empty((n, T, v.A), **f32)

  def one_step():
    v.reset_dev()                                   # collect.py:33 resets before every rollout
    v.rollout_dev(T, fs, bits, act)

  def barrier():
    if world_size > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(a.warmup):
    one_step()
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  launches0 = v.kernel_launches
  v.enable_timing(True)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  kern_ms = []
  barrier()
  e0.record()
  for _ in range(a.steps):
    one_step()
    kern_ms.append(None)
  e1.record()
  launches = v.kernel_launches - launches0     # kernels of this library launched inside the timed region: (k_reset + k_rollout) per step
  barrier()
  # the rollout kernel's own duration (CUDA events recorded by the library around the launch, on the launching stream)
  v.enable_timing(True)
  one_kernel = []
  for _ in range(max(1, min(a.steps, 2))):
    v.reset_dev()
    v.rollout_dev(T, fs, bits, act)
    torch.cuda.synchronize()
    one_kernel.append(v.last_step_ms())
  clocks = sampler.stop() if rank == 0 else None
  total_ms = e0.elapsed_time(e1)
  t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
  if world_size > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  total_ms = float(t.item())
  ms_per_step = total_ms / a.steps
  value = world_size * n * T / (ms_per_step / 1e3)
  overflow = int(v.counters()[:, 5].sum())

  # ---- e2e: vector-env step through host buffers (blcd_step_host), rank-local, then max over ranks ------------------
  e2e = None
  if not a.no_e2e:
    ne = a.e2e_worlds or n
    ve = v if ne == n else VecWorldEnv(env, ne, device=dev, seed=0, world_offset=rank * ne)
    import ctypes as C
    from boxlcd_b200 import _lib
    # the same workload as the device-resident leg: a fresh U[-1, 1) action for every world and step (collect.py:35), drawn on
    # the host beforehand; step t copies its own [N, A] slice host->device
    h_acts = np.random.default_rng(rank).uniform(-1, 1, (max(T, 4), ne, ve.A)).astype(np.float32)
    h_act = h_acts[0]
    ve.pin_host(h_acts)
    h_fs = np.zeros((ne, ve.S), np.float32)
    h_bits = np.zeros((ne,) + ve.bits_shape(), np.uint32)
    h_done = np.zeros(ne, np.uint8)
    Te = T   # the same workload as the device-resident leg: a reset and one full episode of random actions
    # (a) one synchronous call per env step (AsyncVectorEnv.step call shape)
    def e2e_pass(k):
      ve.reset_dev()
      for i in range(k):
        _lib.check(ve.l.blcd_step_host(ve.h, h_acts[i].ctypes.data, h_fs.ctypes.data, h_bits.ctypes.data, h_done.ctypes.data))
    # (b) step_async / step_wait call shape with two steps in flight: a collector's actions do not depend on the step just
    # submitted (collect.py:35), so the copies and the kernel tail of step t hide behind the kernel of step t+1.  Every
    # step still moves its own actions host->device and its own observations device->host; four output slots are cycled.
    ring = [(np.zeros_like(h_fs), np.zeros_like(h_bits), np.zeros_like(h_done)) for _ in range(4)]
    ve.pin_host(*[b for r in ring for b in r])
    def e2e_pass_async(k):
      ve.reset_dev()
      for i in range(k):
        f, b, d = ring[i % 4]
        ve.step_host_async(h_acts[i], f, b, d)
        ve.step_host_wait(keep_in_flight=2)
      ve.step_host_wait(0)
    def timed(fn):
      fn(4)
      barrier()
      t0 = time.perf_counter()
      fn(Te)
      torch.cuda.synchronize()
      tt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
      if world_size > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
      return world_size * ne * Te / float(tt.item())
    e_sync, e_async = timed(e2e_pass), timed(e2e_pass_async)
    e2e = {'value': e_sync, 'unit': 'env-steps/s', 'h2d_bytes_per_step': int(h_act.nbytes),
           'd2h_bytes_per_step': int(h_fs.nbytes + h_bits.nbytes + h_done.nbytes), 'worlds': ne, 'env_steps_timed': Te,
           'api': 'blcd_step_host: host actions in, host full_state + packed frames + done out, one blocking call per env step; '
                  'timed: reset + one full episode of fresh random actions',
           'pipelined_value': e_async, 'pipelined_api': 'blcd_step_host_async + blcd_step_host_wait(keep_in_flight=2), same copies per step'}

  # ---- the rasterizer alone (blcd_render_poses): frames/s and HBM GB/s from poses resident in HBM -------------------------
  render = None
  if rank == 0:
    nr = 4 * 1024 * 1024
    poses, _ = v.get_poses_dev()
    poses = poses[torch.randint(0, n, (nr,), device=dev)].contiguous()
    for _ in range(3):
      v.render_poses_dev(poses)
    torch.cuda.synchronize()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(5):
      v.render_poses_dev(poses)
    r1.record()
    torch.cuda.synchronize()
    r_ms = r0.elapsed_time(r1) / 5
    frame_words = int(np.prod(v.bits_shape()))
    r_bytes = nr * (16 * v.B + 4 * frame_words)
    render = {'kernel': 'k_render_poses', 'frames': nr, 'ms': r_ms, 'frames_per_s': nr / (r_ms / 1e3), 'algorithmic_bytes_per_frame': 16 * v.B + 4 * frame_words,
              'achieved_gbs': r_bytes / (r_ms / 1e3) / 1e9}
    del poses
  if rank != 0:
    if world_size > 1:
      dist.destroy_process_group()
    return
  peaks = {}
  pk_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(pk_path):
    peaks = json.load(open(pk_path))
  peak_gbs, peak_src = (peaks['hbm_gbs'], 'measured') if 'hbm_gbs' in peaks else (6650.0, 'fallback')
  k_ms = float(np.mean(one_kernel))
  alg = ALG_BYTES(sp) * n * T
  traffic, issue = None, None
  prof_path = os.path.join(ROOT, 'profiles', 'rollout_kernel_ncu.json')
  if os.path.exists(prof_path):
    pj = json.load(open(prof_path))
    traffic, issue = pj.get('dram_bytes_per_launch_at_bench_size'), pj.get('issue')
  roofline = {'bound': 'hbm', 'achieved': alg / (k_ms / 1e3) / 1e9, 'peak': peak_gbs, 'unit': 'GB/s', 'frac': alg / (k_ms / 1e3) / 1e9 / peak_gbs,
              'traffic': traffic, 'peak_source': peak_src, 'kernel': 'k_rollout', 'kernel_ms': k_ms, 'algorithmic_bytes_per_env_step': ALG_BYTES(sp),
              'note': 'the fused step is bound by SM issue / dependent fp32 latency of the sequential-impulse solver, not by HBM (SURVEY 8d); '
                      'issue-slot utilisation from ncu is under "issue"', 'issue': issue}
  cpu = None
  if not a.no_cpu:
    cores = host_cores()
    rate, sample = cpu_rollout_rate(sp, T, cores, a.cpu_seconds)
    cpu = {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample}
  line = {
      'metric': 'env-steps/sec incl. LCD frames', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world_size, 'steps': a.steps, 'warmup': a.warmup,
      'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': f'envs.{a.env}() {sp.lcd_h}x{sp.lcd_w}, {n} worlds per GPU, random-action {T}-step rollouts (reset + step + obs + frame)',
                 'worlds_per_gpu': n, 'T': T, 'l2': 'inputs larger than L2 (per-world state + outputs >> 126 MB)', 'scene': info,
                 'solver': 'Box2D 2.3 semantics: 3 sub-steps x (180 velocity + <=60 position iterations), TOI vs walls, sleeping',
                 'manifold_slot_overflows': overflow},
      'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu,
      'render_roofline': None if render is None else dict(render, bound='hbm', peak=peak_gbs, unit='GB/s', frac=render['achieved_gbs'] / peak_gbs),
  }
  print(json.dumps(line), flush=True)
  if world_size > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
