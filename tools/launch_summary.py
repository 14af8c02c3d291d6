"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X`): launches, total time and
share per kernel.  Cold-cache, serialised times: compare SHARES, not absolute numbers."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if 'Kernel Name' in r)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or len(r) <= iv or r[ik] == 'Kernel Name':
        continue
    try:
        v = float(r[iv].replace(',', ''))
    except ValueError:
        continue
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(r[iu], 1e-6)
    t = tot[r[ik]]
    t[0] += 1
    t[1] += v * scale
total = sum(t[1] for t in tot.values())
print(f'# {sys.argv[2] if len(sys.argv) > 2 else ""}')
print('# kernel, launches, total_ms, share')
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
    name = k.split('(')[0][-80:]
    print(f'{name}, {n}, {ms:.4f}, {100 * ms / total:.1f}%')
