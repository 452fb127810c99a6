"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one step's window -> per-kernel totals."""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if r]
hdr_i = next(i for i, r in enumerate(rows) if r[0] == 'ID')
hdr = rows[hdr_i]
ni, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
ui = hdr.index('Metric Unit')
launch = [(r[ni], float(r[vi].replace(',', '')) * (1e-3 if r[ui] in ('ns', 'nsecond') else 1.0)) for r in rows[hdr_i + 1:] if len(r) > vi]
# one step = from one gate forward kernel to the next
starts = [i for i, (n, _) in enumerate(launch) if 'binary_gumbel_fwd' in n or 'hard_concrete_fwd' in n]
print('launches', len(launch), 'step starts at', starts)
if len(starts) >= 2:
    launch = launch[starts[0]:starts[1]]
tot = sum(t for _, t in launch)
agg = OrderedDict()
for n, t in launch:
    key = n.split('(')[0].replace('void ', '').replace('topo::<unnamed>::', '').replace('topo::', '')[:60]
    c, s = agg.get(key, (0, 0.0))
    agg[key] = (c + 1, s + t)
print(f'window: {len(launch)} launches, {tot:.1f} us of kernel time (serialised, cold cache: read the shares)')
print('| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|')
for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{k}` | {c} | {s:.1f} | {s / tot:.3f} | {s / c:.1f} |')
