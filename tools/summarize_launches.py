"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals per step."""
import collections
import csv
import re
import sys

path, steps = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0, []])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
    agg[name][0] += 1; agg[name][1] += v; agg[name][2].append(v); tot += v
print(f"total {tot:.1f} us over {steps:g} steps -> {tot / steps:.1f} us/step, {sum(a[0] for a in agg.values()) / steps:.0f} launches/step")
for k, (n, t, l) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(f"{t / steps:9.1f} us/step {100 * t / tot:5.1f}%  n/step={n / steps:6.1f} avg={t / n:7.1f} min={min(l):7.1f} max={max(l):7.1f}  {k}")
