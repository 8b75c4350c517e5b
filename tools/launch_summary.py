"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/launch_summary.py launches.csv [top]"""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
agg = collections.OrderedDict()
for r in csv.reader(open(path, errors='replace')):
    if r and r[0] == 'ID':
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(d['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(d['Metric Unit'], 1e-6)
    name = d['Kernel Name'].split('(')[0].replace('void ', '').replace('upcp::', '')
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f'{len(agg)} kernels, {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms (cold, serialised)')
for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{100 * ms / tot:5.1f}%  {ms:9.3f} ms  x{cnt:<4d} {name[:90]}')
