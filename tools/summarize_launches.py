#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total us, share.
usage: python tools/summarize_launches.py launches.csv [top_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
by, tot = collections.OrderedDict(), 0.0
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    v = float(r[ival].replace(",", ""))
    v = v / 1000 if r[iunit] in ("ns", "nsecond") else (v * 1000 if r[iunit] in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", r[iname]).replace("void ", "")
    name = re.sub(r"<.*", "", name) if not name.startswith("bhsr::") else name
    e = by.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += v
    tot += v
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ours = sum(v[1] for k, v in by.items() if k.startswith("bhsr::"))
print(f"{sum(v[0] for v in by.values())} launches, {tot / 1e3:.2f} ms of kernels; bhsr:: kernels {ours / 1e3:.2f} ms ({ours / tot:.0%}), "
      f"library / framework kernels {(tot - ours) / 1e3:.2f} ms")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"| `{k[:110]}` | {v[0]} | {v[1]:.0f} | {v[1] / tot:.1%} |")
