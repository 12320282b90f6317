"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel name."""
import csv, re, sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum': continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'nsecond')
    ns = v * {'nsecond': 1, 'ns': 1, 'usecond': 1e3, 'us': 1e3, 'msecond': 1e6, 'ms': 1e6, 'second': 1e9}.get(unit, 1)
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'^.*::', '', name) if '<unnamed>' in r['Kernel Name'] or '_GLOBAL__N' in r['Kernel Name'] else name
    rows.append((name, ns))
tot = sum(ns for _, ns in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1; agg[n][1] += ns
print(f'{len(rows)} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare shares, not absolutes)')
print(f'{"kernel":60s} {"launches":>8s} {"total us":>12s} {"avg us":>10s} {"share":>7s}')
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{n[:60]:60s} {c:8d} {ns / 1e3:12.1f} {ns / 1e3 / c:10.1f} {100 * ns / tot:6.1f}%')
