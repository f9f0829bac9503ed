"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name:
   python tools/agg_launches.py profiles/r01_launches_eager_step_v6.csv [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h, start = r, i
        break
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    n = re.sub(r'\(.*', '', r[ki])[:72]
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print('total %.1f us over %d launches' % (tot / 1e3, sum(v[0] for v in agg.values())))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print('%-74s %5d %9.1f us %5.1f%%' % (n, c, t / 1e3, 100 * t / tot))
