"""Summarise an `ncu --csv --metrics ...` capture: one line per kernel launch (or per kernel name with --sum)."""
import collections, csv, io, sys
NAMES = ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
         'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
         'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread']
t = open(sys.argv[1]).read()
t = t[t.find('"ID"'):]
by = collections.OrderedDict()
for r in csv.DictReader(io.StringIO(t)):
  k = (r['ID'], r['Kernel Name'].split('(')[0].split('::')[-1])
  by.setdefault(k, {})[r['Metric Name']] = (r['Metric Value'], r['Metric Unit'])
first, last = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10**9)
print('kernel        time   issue%  warps%  lanes  warp-inst  dramR  dramW  l1hit%  l2hit%  regs')
tot = 0.0
for i, (k, v) in enumerate(by.items()):
  if not first <= i < last:
    continue
  print(f'{k[1][:13]:13s}', *[' '.join(v.get(n, ('-', ''))) for n in NAMES])
