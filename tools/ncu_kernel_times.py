"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.  python tools/ncu_kernel_times.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(",", ""))
        d[r[ki][:70]].append(v / 1e3 if r[ui] in ("ns", "nsecond") else v)
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v):10.1f} us {100 * sum(v) / tot:5.1f}%  n={len(v):5d} avg {sum(v) / len(v):8.1f} us  {k}")
