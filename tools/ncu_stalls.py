"""Bucket the per-instruction stall samples of one kernel in an .ncu-rep (source page) by code region.
python tools/ncu_stalls.py rep kernel_regex [bucket]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h) and r[2].isdigit()]
# several captured launches of the same kernel are concatenated: keep the first
addr0 = data[0][0]
rep_at = [i for i, r in enumerate(data) if r[0] == addr0]
if len(rep_at) > 1:
    data = data[:rep_at[1]]
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
ie = h.index("Instructions Executed")
print("instructions", len(data), "samples", sum(int(r[2]) for r in data))
for b in range(0, len(data), bucket):
    seg = data[b:b + bucket]
    s = sum(int(r[2]) for r in seg)
    if s < 50:
        continue
    agg = {}
    for r in seg:
        for st in stalls:
            v = int(r[h.index(st)] or 0)
            if v:
                agg[st[6:]] = agg.get(st[6:], 0) + v
    ex = sum(int(r[ie]) for r in seg)
    top = max(seg, key=lambda r: int(r[2]))
    print(f"{b:5d} samples {s:6d} exec {ex:9d} {sorted(agg.items(), key=lambda x: -x[1])[:4]} | hot: {top[2]} {top[1][:60]}")
