"""Per-CUDA-source-line stall samples of one launch in an .ncu-rep (needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys

rep, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# sections start with "File Path"; pick the `which`-th kernel's sections
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'File Path':
        cur = {'file': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
# group blocks into launches: a launch restarts when the same file path appears again
launches, seen = [], set()
for b in blocks:
    if b['file'] in seen:
        seen = set()
    if not seen:
        launches.append([])
    seen.add(b['file'])
    launches[-1].append(b)
total = 0
lines = []
for b in launches[which]:
    hdr = None
    for r in b['rows']:
        if r and r[0] == 'Line No':
            hdr = r
            si = hdr.index('# Samples')
            ex = hdr.index('Instructions Executed')
            continue
        if hdr is None or len(r) <= si or not r[0].isdigit():
            continue
        n = int(r[si]) if r[si].isdigit() else 0
        total += n
        lines.append((b['file'].split('/')[-1], int(r[0]), n, r[ex], r[1]))
print('total samples', total)
for f, ln, n, ex, src in lines:
    if n >= total * thresh / 100:
        print(f'{f:22s} {ln:5d} {n:6d} {100.0 * n / total:5.1f}% exec={ex:>9s}  {src.strip()[:110]}')
