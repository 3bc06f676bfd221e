"""bench.py -- env-steps/sec (each step includes one LCD frame, full_state and the action) for the B200-native boxLCD
hot path, next to the CPU path on the box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--env Urchin] [--worlds 262144] [--scaling strong|weak] [--T 100]
                  [--impl reference] [--sweep]

A "step" is one pass of the hot path over one batch: every world of this rank's shard is reset and runs a T-env-step
random-action rollout (examples/collect.py:31-39) on the device (blcd_reset + blcd_rollout: action RNG -> WorldEnv.step ->
_get_obs -> lcd_render, dataset rows written straight to HBM).  value = env-steps of all ranks / max-over-ranks device time.
e2e = the same count through blcd_step_host (host action buffer in, host observation buffers out, every env step).

Scaling ("262 144 worlds at 1/2/4/8 GPUs, N split evenly"): `--worlds` is the TOTAL world
count and is split evenly over the ranks ("scaling": "strong", the default).  `--scaling weak` gives every rank `--worlds`
worlds instead.  No collective on the data path: worlds are independent and keyed by global index.

--sweep: BASELINE config 5 -- UrchinBall at 4 096 ... 1 048 576 TOTAL worlds over the N ranks, one JSON line with the table.
"""
import argparse
import csv
import io
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
BASELINE_T = {'Dropbox': 200}     # `--ep_len=200`; every other config uses the env's ep_len
SWEEP_WORLDS = [4096, 16384, 65536, 262144, 1048576]

# fp32 operations per constraint solve, counted from the restatement's source (oracle/b2_world.h; the CUDA path performs the
# same arithmetic): revolute velocity solve (motor 11 + Cdot 10 + 3x3 solve 40 + 2x2 solve 12 + apply 22), contact point
# velocity solve (tangent 36 + normal 34; the 2-point block solver costs about the same per point), revolute position solve
# (limit 15 + two sincos 40 + anchors / error 30 + 2x2 mass and solve 30 + apply 18), contact point position solve (two
# sincos 40 + transforms 20 + separation / impulse 25 + apply 18), narrow phase per candidate pair (AABB test + manifold
# ~150), integration / damping / clamps per body 30.  SURVEY 8(d) formula.
F_JOINT_VEL, F_CONTACT_VEL, F_JOINT_POS, F_CONTACT_POS, F_PAIR, F_BODY = 95, 70, 133, 103, 150, 30


def parse():
  p = argparse.ArgumentParser()
  p.add_argument('--gpus', type=int, default=1)
  p.add_argument('--steps', type=int, default=3)
  p.add_argument('--warmup', type=int, default=3)
  p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  p.add_argument('--env', default='Urchin')
  p.add_argument('--worlds', type=int, default=262144, help='TOTAL worlds (split evenly over the ranks); per GPU with --scaling weak')
  p.add_argument('--scaling', default='strong', choices=['strong', 'weak'])
  p.add_argument('--T', type=int, default=0, help="env steps per rollout (0 = BASELINE's figure for the env: its ep_len, 200 for Dropbox)")
  p.add_argument('--e2e_worlds', type=int, default=0, help='worlds per rank for the host-buffer e2e leg (0 = same as the device leg)')
  p.add_argument('--cpu_seconds', type=float, default=12.0, help='target CPU work for the cpu_baseline sample')
  p.add_argument('--sweep', action='store_true', help='BASELINE config 5: UrchinBall, 4 096 ... 1 048 576 total worlds')
  p.add_argument('--no_cpu', action='store_true')
  p.add_argument('--no_e2e', action='store_true')
  p.add_argument('--no_ncu', action='store_true', help='skip the ncu sub-process that counts the instructions of one rollout')
  p.add_argument('--no_render', action='store_true')
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


def shard(total, rank, world):
  return total * rank // world, total * (rank + 1) // world


def cpu_rollout_rate(spec, T, cores, seconds):
  """times the CPU oracle (oracle/b2_oracle.cpp: the restatement of pybox2d + PIL this repo checks against, kind
  "port") on `cores` threads over a bounded sample of the same workload: reset + T-step random-action rollouts with
  frames.  Returns (env-steps/s, sample description, mean diagnostic counters per sub-step)."""
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
  c = ow.counters().astype(np.float64).sum(0)
  names = oracle.COUNTER_NAMES
  sub = max(c[names.index('substeps')], 1.0)
  per_sub = {k: float(c[names.index(k)] / sub) for k in ('contacts', 'pos_iters', 'toi_events', 'toi_calls', 'manifold_points')}
  return n * T / dt, f'{n} worlds x {T} env-steps (reset + rollout + frames), {dt:.1f} s on {cores} threads', per_sub


def flops_per_env_step(sp, per_sub, n_pairs):
  """3 x [ VI (J F_j + C F_c) + PI (J G_j + C G_c) + pairs F_np + B F_int ] with the iteration / contact counts
  measured on the oracle in this run (per_sub: means per b2World.Step) and the per-solve constants above."""
  J, B = sp.n_joints, sp.n_bodies
  C = per_sub['manifold_points']          # contact POINTS per sub-step (each is one normal + one tangent solve per sweep)
  PI = per_sub['pos_iters']               # position sweeps per sub-step (islands that were still iterating)
  per_sub_flops = sp.vel_iters * (J * F_JOINT_VEL + C * F_CONTACT_VEL) + PI * (J * F_JOINT_POS + C * F_CONTACT_POS) + n_pairs * F_PAIR + B * F_BODY
  return sp.n_substeps * per_sub_flops


def run_reference(a, rank, world_size):
  """--impl reference: the reference's CPU path for the same workload.  pybox2d is not installable in this image, so
  the arm runs the repo's CPU restatement of it (oracle/, kind "port") on all host threads -- never presented as pybox2d."""
  if rank != 0:
    return
  import boxlcd_b200 as blcd
  env = blcd.env_map[a.env]()
  sp = env.layout.spec
  T = a.T or BASELINE_T.get(a.env, env.G.ep_len)
  cores = host_cores()
  from oracle import oracle
  total = a.worlds * (world_size if a.scaling == 'weak' else 1)
  per_step = max(cores, int(cores * 9000 * 6.0 / T) // cores * cores)   # ~6 s of CPU work per step on ~9k steps/s/core
  per_step = min(per_step, total)
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
      'warmup': a.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': a.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': workload_name(a.env, sp, total, T), 'total_worlds': total, 'sample_worlds_per_step': per_step,
                 'sample_fraction': per_step / total, 'T': T,
                 'note': 'CPU restatement of Box2D 2.3 + PIL rasterizer (oracle/), not pybox2d: pybox2d is not installable here; each step times a '
                         'bounded sample of the workload (sample_fraction of its worlds), the rate is per env-step and does not depend on the sample size'},
      'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': f'{per_step} worlds x {T} env-steps per step'},
      'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(line), flush=True)


def workload_name(env_name, sp, total, T):
  return f'envs.{env_name}() {sp.lcd_h}x{sp.lcd_w}, {total} worlds in total, random-action {T}-step rollouts (reset + step + obs + frame)'


def ncu_instruction_count(env_name, n, T, pipeline, timeout=300):
  """Count, IN THIS RUN, the instructions of one rollout: an `ncu --metrics` sub-process profiles every kernel this library
  launches for one T-step rollout of n worlds (tools/ncu_case.py brackets it with cudaProfilerStart/Stop) and the
  per-kernel counters are summed.  Returns None when ncu is unavailable (then nothing is reported -- no static copy)."""
  metrics = ('smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,'
             'dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum')
  cmd = ['ncu', '--csv', '--profile-from-start', 'off', '--clock-control', 'none', '--metrics', metrics,
         sys.executable, os.path.join(ROOT, 'tools', 'ncu_case.py'), env_name, str(n), str(T), 'range']
  try:
    # the sub-process must take the same device path as the timed handle (the library picks it from the world count)
    # ... and one world range per launch: under ncu the kernels are serialised anyway, and every launch then covers all worlds
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=dict(os.environ, BLCD_PIPELINE=str(int(pipeline)), BLCD_PIPE_RANGES='1'))
  except Exception as e:
    return {'unavailable': f'{type(e).__name__}: {e}'[:200]}
  text = res.stdout
  start = text.find('"ID"')
  if res.returncode != 0 or start < 0:
    return {'unavailable': (res.stderr or res.stdout)[-300:].replace('\n', ' ')}
  tot, kernels, launches = {}, {}, {}
  for row in csv.DictReader(io.StringIO(text[start:])):
    try:
      val = float(row['Metric Value'].replace(',', ''))
    except Exception:
      continue
    unit = row.get('Metric Unit', '')
    k = row['Kernel Name'].split('(')[0].split('::')[-1]
    if row['Metric Name'] == 'gpu__time_duration.sum':
      val *= {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}.get(unit, 1e-9)
      kernels[k] = kernels.get(k, 0.0) + val
    if row['Metric Name'].startswith('dram__bytes'):
      val *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    launches.setdefault((row['ID'], k), {})[row['Metric Name']] = val
    if not row['Metric Name'].startswith('smsp__issue_active'):
      tot[row['Metric Name']] = tot.get(row['Metric Name'], 0.0) + val
  # per kernel: share of the GPU time, lanes per instruction, issue-slot utilisation (duration-weighted over its launches)
  per_kernel = {}
  for (_, k), m in launches.items():
    a = per_kernel.setdefault(k, {'seconds': 0.0, 'warp_inst': 0.0, 'lane_inst': 0.0, 'issue_x_s': 0.0, 'launches': 0})
    dur = m.get('gpu__time_duration.sum', 0.0)
    a['seconds'] += dur; a['launches'] += 1
    a['warp_inst'] += m.get('smsp__inst_executed.sum', 0.0); a['lane_inst'] += m.get('smsp__thread_inst_executed.sum', 0.0)
    a['issue_x_s'] += m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0.0) * dur
  all_s = max(sum(a['seconds'] for a in per_kernel.values()), 1e-12)
  for k, a in per_kernel.items():
    a['share_of_gpu_time'] = a['seconds'] / all_s
    a['active_lanes_per_warp_inst'] = a['lane_inst'] / max(a['warp_inst'], 1.0)
    a['issue_slot_utilisation_pct'] = a.pop('issue_x_s') / max(a['seconds'], 1e-12)
    a['frac_of_lane_issue_peak'] = a['issue_slot_utilisation_pct'] / 100.0 * a['active_lanes_per_warp_inst'] / 32.0
  if 'smsp__thread_inst_executed.sum' not in tot:
    return {'unavailable': 'ncu printed no counters'}
  steps = float(n) * T
  return {'worlds': n, 'T': T, 'lane_inst_per_env_step': tot['smsp__thread_inst_executed.sum'] / steps,
          'warp_inst_per_env_step': tot['smsp__inst_executed.sum'] / steps,
          'active_lanes_per_warp_inst': tot['smsp__thread_inst_executed.sum'] / max(tot['smsp__inst_executed.sum'], 1.0),
          'dram_bytes_per_env_step': (tot.get('dram__bytes_read.sum', 0.0) + tot.get('dram__bytes_write.sum', 0.0)) / steps,
          'per_kernel_under_ncu': dict(sorted(per_kernel.items(), key=lambda kv: -kv[1]['seconds'])[:8]),
          'source': 'ncu sub-process of this bench run (cold caches, serialised kernels): counts, not times, are used'}


def main():
  a = parse()
  rank = int(os.environ.get('RANK', 0))
  world_size = int(os.environ.get('WORLD_SIZE', 1))
  local_rank = int(os.environ.get('LOCAL_RANK', 0))
  if a.sweep:
    a.env = 'UrchinBall'
  if a.impl == 'reference':
    run_reference(a, rank, world_size)
    return
  import torch
  import torch.distributed as dist
  import boxlcd_b200 as blcd
  from boxlcd_b200 import _lib
  from boxlcd_b200.vec_env import VecWorldEnv
  assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback in the product path)'
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world_size > 1:
    dist.init_process_group('nccl', device_id=dev)
  env = blcd.env_map[a.env]()
  sp = env.layout.spec
  T = a.T or BASELINE_T.get(a.env, env.G.ep_len)

  def barrier():
    if world_size > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(x):
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world_size > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  def time_rollouts(total, steps, warmup):
    """(ms per step, max over ranks; launches per step; VecWorldEnv; output tensors) for `total` worlds split over the ranks"""
    lo, hi = shard(total, rank, world_size) if a.scaling == 'strong' else (rank * total, (rank + 1) * total)
    n = hi - lo
    v = VecWorldEnv(env, max(n, 1), device=dev, seed=0, world_offset=lo)
    f32 = dict(dtype=torch.float32, device=dev)
    fs = torch.empty((v.n, T, v.S), **f32)
    bits = torch.empty((v.n, T) + v.bits_shape(), dtype=torch.int32, device=dev)
    act = torch.empty((v.n, T, v.A), **f32)

    def one_step():
      v.reset_dev()                                   # collect.py:33 resets before every rollout
      v.rollout_dev(T, fs, bits, act)
    for _ in range(warmup):
      one_step()
    barrier()
    l0 = v.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      one_step()
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)) / steps, (v.kernel_launches - l0), v, (fs, bits, act), one_step

  total = a.worlds if a.scaling == 'strong' else a.worlds * world_size

  # ---- config 5: world-count sweep ------------------------------------------------------------------------------------
  if a.sweep:
    rows = []
    for tw in SWEEP_WORLDS:
      ms, _, v, bufs, _ = time_rollouts(tw if a.scaling == 'strong' else tw // world_size, max(1, a.steps), max(1, min(a.warmup, 2)))
      tw_eff = tw
      rows.append({'total_worlds': tw_eff, 'worlds_per_gpu': v.n, 'ms_per_step': ms, 'env_steps_per_s': tw_eff * T / (ms / 1e3), 'block': v.info()['block']})
      v.close()
      del v, bufs
      torch.cuda.empty_cache()
    if rank == 0:
      cpu = None
      if not a.no_cpu:
        cores = host_cores()
        rate, sample, _ = cpu_rollout_rate(sp, T, cores, a.cpu_seconds)
        cpu = {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample}
      best = max(rows, key=lambda r: r['env_steps_per_s'])
      print(json.dumps({'metric': 'env-steps/sec incl. LCD frames', 'value': best['env_steps_per_s'], 'unit': 'env-steps/s', 'n_gpus': world_size,
                        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': best['ms_per_step'], 'higher_is_better': True, 'scaling': a.scaling,
                        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                        'config': {'workload': f'envs.UrchinBall() 16x24 world-count sweep, {T}-step random-action rollouts (BASELINE config 5); value = best row',
                                   'T': T, 'l2': 'inputs larger than L2 from 65 536 worlds up'},
                        'sweep': rows, 'cpu_baseline': cpu}), flush=True)
    if world_size > 1:
      dist.destroy_process_group()
    return

  # ---- the headline leg: device-resident rollouts -----------------------------------------------------------------------
  sampler = ClockSampler(local_rank)
  # warm-up happens inside time_rollouts; clocks are sampled over warm-up + timed region + the per-launch timing below
  if rank == 0:
    sampler.start()
  ms_per_step, launches, v, (fs, bits, act), one_step = time_rollouts(a.worlds, a.steps, a.warmup)
  n = v.n
  info = v.info()
  value = total * T / (ms_per_step / 1e3)
  # the rollout's own duration (CUDA events recorded by the library around the launch(es), on the launching stream)
  v.enable_timing(True)
  one_kernel = []
  for _ in range(max(1, min(a.steps, 2))):
    v.reset_dev()
    v.rollout_dev(T, fs, bits, act)
    torch.cuda.synchronize()
    one_kernel.append(v.last_step_ms())
  v.enable_timing(False)
  clocks = sampler.stop() if rank == 0 else None
  overflow = int(v.counters()[:, 5].sum())

  # ---- e2e: vector-env step through host buffers (blcd_step_host), rank-local, then max over ranks ------------------
  e2e = None
  if not a.no_e2e:
    ne = a.e2e_worlds or n
    ve = v if ne == n else VecWorldEnv(env, ne, device=dev, seed=0, world_offset=rank * ne)
    # the same workload as the device-resident leg: a fresh U[-1, 1) action for every world and step (collect.py:35), drawn on
    # the host beforehand; step t copies its own [N, A] slice host->device.  The caller's buffers are page-locked once with
    # blcd_pin_host (they live for the whole run); blcd_step_host then copies straight from / into them.
    h_acts = np.random.default_rng(rank).uniform(-1, 1, (max(T, 4), ne, ve.A)).astype(np.float32)
    h_fs = np.zeros((ne, ve.S), np.float32)
    h_bits = np.zeros((ne,) + ve.bits_shape(), np.uint32)
    h_done = np.zeros(ne, np.uint8)
    ring = [(np.zeros_like(h_fs), np.zeros_like(h_bits), np.zeros_like(h_done)) for _ in range(4)]
    ve.pin_host(h_acts, h_fs, h_bits, h_done, *[b for r in ring for b in r])
    Te = T

    def e2e_pass(k):     # (a) one synchronous call per env step (AsyncVectorEnv.step call shape)
      ve.reset_dev()
      for i in range(k):
        _lib.check(ve.l.blcd_step_host(ve.h, h_acts[i].ctypes.data, h_fs.ctypes.data, h_bits.ctypes.data, h_done.ctypes.data))

    def e2e_pass_async(k):   # (b) step_async / step_wait call shape with two steps in flight; same copies per step
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
      dt = max_over_ranks(time.perf_counter() - t0)
      return world_size * ne * Te / dt
    e_sync, e_async = timed(e2e_pass), timed(e2e_pass_async)
    e2e = {'value': e_sync, 'unit': 'env-steps/s', 'h2d_bytes_per_step': int(h_acts[0].nbytes),
           'd2h_bytes_per_step': int(h_fs.nbytes + h_bits.nbytes + h_done.nbytes), 'worlds_per_gpu': ne, 'env_steps_timed': Te,
           'api': 'blcd_step_host: host actions in, host full_state + packed frames + done out, one blocking call per env step; '
                  'timed: reset + one full episode of fresh random actions',
           'pipelined_value': e_async, 'pipelined_api': 'blcd_step_host_async + blcd_step_host_wait(keep_in_flight=2), same copies per step'}

  # ---- the rasterizer alone (blcd_render_poses): frames/s and HBM GB/s from poses resident in HBM -------------------------
  render = None
  if rank == 0 and not a.no_render:
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
    render = {'kernel': 'k_render_bodies', 'frames': nr, 'ms': r_ms, 'frames_per_s': nr / (r_ms / 1e3), 'algorithmic_bytes_per_frame': 16 * v.B + 4 * frame_words,
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
  cpu, per_sub = None, None
  if not a.no_cpu:
    cores = host_cores()
    rate, sample, per_sub = cpu_rollout_rate(sp, T, cores, a.cpu_seconds)
    cpu = {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample, 'counters_per_substep': per_sub}

  # ---- solver roofline: everything measured in this run ------------------------------------------------------------------
  import ctypes as C
  pk = (C.c_double * 4)()
  _lib.check(v.l.blcd_measure_peaks(local_rank, pk))
  lane_peak, chain_rate = pk[0], pk[1]
  counts = None
  if not a.no_ncu:
    torch.cuda.synchronize()
    counts = ncu_instruction_count(a.env, n, 1 if info.get('pipeline', 0) else 3, info.get('pipeline', 0))
  per_gpu_rate = n * T / (k_ms / 1e3)      # env-steps/s of this GPU inside the rollout launch(es)
  solver = {'bound': 'issue', 'unit': 'lane-instructions/s', 'peak': lane_peak,
            'peak_source': 'blcd_measure_peaks in this run: fp32 FMA with every issue slot filled (x2 = %.1f TFLOP/s fp32)' % (2 * lane_peak / 1e12),
            'dependent_chain_ffma_per_s': chain_rate,
            'note': 'achieved = lane-instructions per env-step (ncu counters of one rollout, this run) x env-steps/s of the timed rollout; '
                    'a sequential-impulse sweep is a serial fp32 recurrence per world, so the attainable fraction is bounded by occupancy x lanes kept busy'}
  if counts and 'lane_inst_per_env_step' in counts:
    solver.update(achieved=counts['lane_inst_per_env_step'] * per_gpu_rate, frac=counts['lane_inst_per_env_step'] * per_gpu_rate / lane_peak,
                  issue_slot_utilisation=counts['warp_inst_per_env_step'] * per_gpu_rate * 32 / lane_peak, counters=counts)
    top = next(iter(counts['per_kernel_under_ncu'].items()), None)
    if top:
      solver['dominant_kernel'] = dict(top[1], name=top[0], note='share, lanes and issue-slot utilisation of the kernel with the largest share of the GPU time '
                                       '(ncu sub-process of this run); frac_of_lane_issue_peak = issue utilisation x active lanes / 32')
  else:
    solver.update(achieved=None, frac=None, counters=counts)
  if per_sub is not None:
    fl = flops_per_env_step(sp, per_sub, info['n_pairs'])
    solver['flops'] = {'per_env_step': fl, 'achieved_tflops': fl * per_gpu_rate / 1e12, 'peak_tflops': 2 * lane_peak / 1e12,
                       'frac': fl * per_gpu_rate / (2 * lane_peak),
                       'formula': 'n_substeps x [vel_iters (J F_j + C F_c) + PI (J G_j + C G_c) + pairs F_np + B F_int], C and PI measured on the oracle in this run '
                                  '(cpu_baseline.counters_per_substep), per-solve constants counted from the source (bench.py)'}
  traffic = None
  if counts and 'dram_bytes_per_env_step' in counts:
    traffic = counts['dram_bytes_per_env_step'] * n * T
  roofline = {'bound': 'hbm', 'achieved': alg / (k_ms / 1e3) / 1e9, 'peak': peak_gbs, 'unit': 'GB/s', 'frac': alg / (k_ms / 1e3) / 1e9 / peak_gbs,
              'traffic': traffic, 'traffic_source': None if traffic is None else 'dram bytes per env-step from the ncu sub-process of this run x env-steps per launch',
              'peak_source': peak_src, 'kernel': 'blcd_rollout (all kernels of one rollout call)', 'kernel_ms': k_ms, 'algorithmic_bytes_per_env_step': ALG_BYTES(sp),
              'note': 'reported because the contract asks for it: the rollout is bound by SM issue / dependent fp32 latency of the sequential-impulse solver, '
                      'not by HBM (SURVEY 8d) -- see roofline_solver'}
  line = {
      'metric': 'env-steps/sec incl. LCD frames', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world_size, 'steps': a.steps, 'warmup': a.warmup,
      'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': a.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': workload_name(a.env, sp, total, T), 'total_worlds': total, 'worlds_per_gpu': n, 'T': T,
                 'l2': 'inputs larger than L2 (per-world state + outputs >> 126 MB)' if n * T * ALG_BYTES(sp) > 126e6 else 'outputs smaller than L2: see the sweep for sizes',
                 'scene': info, 'solver': 'Box2D 2.3 semantics: 3 sub-steps x (180 velocity + <=60 position iterations), TOI vs walls, sleeping',
                 'manifold_slot_overflows': overflow},
      'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'roofline_solver': solver, 'cpu_baseline': cpu,
      'render_roofline': None if render is None else dict(render, bound='hbm', peak=peak_gbs, unit='GB/s', frac=render['achieved_gbs'] / peak_gbs),
  }
  print(json.dumps(line), flush=True)
  if world_size > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
