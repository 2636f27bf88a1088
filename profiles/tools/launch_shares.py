#!/usr/bin/env python3
"""Per-kernel share of the step from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    agg[r[ki].split('(')[0]].append(v)
tot = sum(sum(v) for v in agg.values())
print(f'{sys.argv[1]}: cold-cache, serialised launches; share of step per kernel')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f'  {k:18s} launches {len(v):3d}  avg {sum(v)/len(v):8.1f} us  share {100*sum(v)/tot:5.1f}%')
