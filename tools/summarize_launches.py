"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = defaultdict(float)
cnt = defaultdict(int)
for i, row in enumerate(r):
    if i < skip:
        continue
    v = float(row[iv].replace(',', ''))
    unit = row[iu]
    ms = v / 1e6 if unit in ('ns', 'nsecond') else v / 1e3 if unit in ('us', 'usecond') else v
    name = row[ik].split('(')[0]
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print(f'{"kernel":60s} {"launches":>8s} {"ms":>10s} {"share":>7s}')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f'{k[:60]:60s} {cnt[k]:8d} {v:10.3f} {100 * v / total:6.1f}%')
print(f'{"TOTAL":60s} {sum(cnt.values()):8d} {total:10.3f}')
