"""Aggregate an ncu report's SASS-level counters by CUDA source function (needs -lineinfo):
   python tools/ncu_by_function.py <report.ncu-rep | prefix of pre-dumped pages> <lib.so> <kernel-substring>
With a prefix P the pages are read from P.raw.csv / P.source.csv (`ncu -i rep --page raw|source --csv`, dumped on the GPU box:
the reports themselves are too large to bring back)."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
# the library holds one cubin per scene-size profile; take the one whose kernel matches (kname may include the profile namespace)
dis = []
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith('.cubin')):
  d = subprocess.run(['nvdisasm', '--print-line-info', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
  if any(l.startswith('.text.') and all(k in l for k in kname.split('+')) for l in d):
    dis = d
    break
def page(name):
  if rep.endswith('.ncu-rep'):
    return subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
  return open(f'{rep}.{name}.csv').read()
raw = page('raw')
rows = list(csv.reader(raw.split('\n')))
hdr, vals = rows[0], rows[2]
get = lambda k: vals[hdr.index(k)]
for k in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
          'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
          'smsp__average_warp_latency_per_inst_issued.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_shared_loads', 'sass__inst_executed_global_loads']:
  if k in hdr: print(f'{k:90s} {get(k)}')
srcrows = list(csv.reader(page('source').split('\n')))
kfull = srcrows[0][1]
start = [i for i, l in enumerate(dis) if l.startswith('.text.') and all(k in l for k in kname.split('+'))][0]
off2loc, cur = {}, ('?', 0)
for l in dis[start + 1:]:
  if l.startswith('//--------------------- .'): break
  m = re.search(r'//## File "([^"]+)", line (\d+)', l)
  if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
  m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
  if m: off2loc[int(m.group(1), 16)] = cur
h = srcrows[1]
ia, ii, it, ist = h.index('Address'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
body = [r for r in srcrows[2:] if len(r) > ist]
base = int(body[0][ia], 16)
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'boxlcd_b200', 'csrc')
src = {f: open(os.path.join(root, f)).read().split('\n') for f in os.listdir(root) if f.endswith(('.cuh', '.cu', '.h'))}
def fn_of(f, ln):
  if f not in src: return f
  for k in range(min(ln, len(src[f])) - 1, -1, -1):
    s = src[f][k]
    m = re.match(r'\s*BLCD_HDN?\s+.*?\b([a-zA-Z_0-9]+)\s*\(', s)
    if m and not s.strip().startswith('//'): return f.split('.')[0] + ':' + m.group(1)
  return f
I, T, S = collections.Counter(), collections.Counter(), collections.Counter()
for r in body:
  k = fn_of(*off2loc.get(int(r[ia], 16) - base, ('?', 0)))
  I[k] += int(r[ii]); T[k] += int(r[it]); S[k] += int(r[ist])
tot, stot = sum(I.values()), sum(S.values())
print(f'kernel {kfull[:80]}  SASS instructions {len(body)}  warp-inst {tot/1e6:.1f} M')
print('%-45s %8s %6s %6s %8s' % ('function (inlined callee lines count for the callee)', 'Minst', '%inst', 'lanes', '%samples'))
for k, n in I.most_common(28):
  print('%-45s %8.1f %6.1f %6.1f %8.1f' % (k, n / 1e6, 100 * n / tot, T[k] / max(n, 1), 100 * S[k] / max(stot, 1)))
