#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
usage: python scripts/launch_summary.py launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1000.0 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1000.0 if u in ("ms", "msecond") else v
    rows.append((r[ki], us))
rows = rows[skip:]
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0.0, 0])
for k, v in rows:
    k = re.sub(r"\(.*", "", k)
    agg[k][0] += v
    agg[k][1] += 1
print(f"# {path}: total {tot:.0f} us over {len(rows)} launches (cold-cache, serialised: compare SHARES)")
for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v:12.1f} us {100*v/tot:5.1f}% n={n:5d} avg {v/n:9.2f} us  {k[:100]}")
