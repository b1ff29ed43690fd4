"""Counts the Blackwell-specific SASS mnemonics per kernel of the built library (evidence that the hot kernels are tcgen05 /
TMEM / TMA code): python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "iad-r1_b200", "libiadr1_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UBLKRED|UBLKCP|UBLKPF|REDG\.E\.ADD\.F32x4|UTCBAR|UTCATOMSWS|HMMA|MOVM|SYNCS|MUFU\.EX2)[\w.]*")
fn, counts = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)", "<anon>").split("(")[0]
        continue
    m = pat.search(line)
    if m and fn:
        counts[fn][m.group(0)] += 1
print("# cuobjdump -sass iad-r1_b200/libiadr1_b200.so : Blackwell mnemonics per kernel (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
print("# UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, UBLKCP/UBLKRED/UBLKPF = bulk copy / bulk reduce / bulk L2 prefetch, REDG.E.ADD.F32x4 = 16-byte vector reduction, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync, MOVM = movmatrix)")
for fn in sorted(counts):
    print(f"{fn}: " + ", ".join(f"{k} x{v}" for k, v in sorted(counts[fn].items())))
