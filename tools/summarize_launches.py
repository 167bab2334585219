"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline='') as f:
    lines = [ln for ln in f if not ln.startswith('==')]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1e-3)
    tot[name][0] += 1
    tot[name][1] += v * scale
total = sum(v[1] for v in tot.values())
print(f'{"kernel":70s} {"launches":>9s} {"total_us":>12s} {"share":>7s} {"avg_us":>9s}')
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{name[:70]:70s} {n:9d} {us:12.1f} {100 * us / total:6.1f}% {us / n:9.2f}')
print(f'{"TOTAL":70s} {sum(v[0] for v in tot.values()):9d} {total:12.1f}')
