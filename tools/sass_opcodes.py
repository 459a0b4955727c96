"""Counts of the Blackwell-specific SASS opcodes per kernel of libax2d.so (evidence that the contraction kernels run on
tcgen05 / TMEM / TMA and the aggregation on bulk async copies + packed / mixed-precision adds).
    python tools/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "aimnet_x2d_b200", "libax2d.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "FADD2", "FFMA2", "FHADD", "LDGSTS", "ACQBULK", "REDUX"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    for op in OPS:
        if re.search(r"\b" + op + r"\b|\b" + op + r"\.", line):
            counts[cur][op] += 1
print("# SASS opcode counts per kernel (cuobjdump -sass aimnet_x2d_b200/libax2d.so); UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,")
print("# UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,")
print("# FADD2 / FFMA2 = add / fma .f32x2, FHADD = add.f32.bf16 (mixed precision), LDGSTS = cp.async, ACQBULK = griddepcontrol.wait, REDUX = warp reduction")
print(f"{'kernel':70s} " + " ".join(f"{o:>8s}" for o in OPS))
tot = collections.Counter()
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    tot.update(c)
    print(f"{k[:70]:70s} " + " ".join(f"{c[o]:8d}" for o in OPS))
print(f"{'TOTAL':70s} " + " ".join(f"{tot[o]:8d}" for o in OPS))
