#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name, share of the total."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "ID":
        hdr = r
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    name = re.sub(r"\(.*", "", r[ki])[:90]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"total {total:.1f} us over {len(rows)} launches")
for name, v in sorted(tot.items(), key=lambda kv: -kv[1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v:10.1f} us {100 * v / total:5.1f}%  x{cnt[name]:<5d} {name}")
