"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py x.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v if r[ui] in ("us", "usecond") else v * 1e3
    name = r[ki].split("(")[0].replace("void ", "").replace("iadr1::", "")[:48]
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
print(f"launches {sum(v[0] for v in tot.values())}, total {total / 1e3:.1f} ms (cold-cache, serialised)")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{100 * us / total:6.2f}%  {us / 1e3:9.2f} ms  {n:6d} x {us / n:9.2f} us  {name}")
