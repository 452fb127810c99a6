"""Top source lines per stall reason for one launch of an .ncu-rep."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
names = ['stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_barrier', 'stall_selected', 'stall_not_selected', 'stall_math', 'stall_lg',
         'stall_mio', 'stall_sleep', 'stall_branch_resolving', 'stall_dispatch', 'stall_no_inst', 'stall_membar', 'stall_tex', 'stall_drain']
hdr, agg, tot, f = None, {}, {}, ''
for r in rows:
    if r and r[0] == 'File Path':
        f = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or not r or not r[0].isdigit():
        continue
    d = {}
    for n in names:
        if n in hdr:
            v = r[hdr.index(n)]
            d[n] = int(v) if v.isdigit() else 0
            tot[n] = tot.get(n, 0) + d[n]
    agg[(f, int(r[0]))] = (d, r[1].strip()[:90])
print({k: v for k, v in tot.items() if v})
for n in sys.argv[2:] or ['stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_barrier']:
    print('==', n, tot.get(n))
    for (ff, ln), (d, src) in sorted(agg.items(), key=lambda kv: -kv[1][0].get(n, 0))[:8]:
        print(f'   {ff}:{ln:4d} {d.get(n, 0):6d}  {src}')
