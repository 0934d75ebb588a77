#!/usr/bin/env python3
"""Summarise an ncu launch-list CSV (gpu__time_duration.sum) per kernel name: launches, total us, share."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
S = sum(v[1] for v in tot.values())
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"| `{k[:110]}` | {n} | {us:.1f} | {us/n:.1f} | {100*us/S:.1f}% |")
print(f"\ntotal {S:.1f} us over {sum(v[0] for v in tot.values())} launches")
