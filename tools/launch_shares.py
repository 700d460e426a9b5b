#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python tools/launch_shares.py file.csv [title]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
if len(sys.argv) > 2:
    print(sys.argv[2])
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4]].append(float(r[-1]) / 1e3)           # ns -> us
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f'{sum(v):12.1f} us {100 * sum(v) / tot:5.1f}%  n={len(v):4d}  avg {sum(v) / len(v):9.1f} us  {k[:110]}')
