"""Per-kernel totals of ONE training step out of an `ncu --graph-profiling node --metrics gpu__time_duration.sum --csv`
launch list of `bench.py --steps 1 --warmup 1`: the step is delimited by two consecutive pack_weights launches."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [re.sub(r"\(.*", "", r["Kernel Name"])[:70] for r in rows]
idx = [i for i, n in enumerate(names) if "pack_weights" in n]
seg = range(idx[-2], idx[-1])
agg = collections.defaultdict(lambda: [0, 0.0, []])
tot = 0.0
for i in seg:
    v = float(rows[i]["Metric Value"].replace(",", ""))
    u = rows[i]["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    agg[names[i]][0] += 1; agg[names[i]][1] += v; agg[names[i]][2].append(v); tot += v
print(f"one step: {tot:.1f} us in {len(seg)} launches (serialised, cold caches: compare shares)")
for k, (n, t, l) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"{t:8.1f} us {100 * t / tot:5.1f}% n={n:3d} avg={t / n:7.1f} min={min(l):6.1f} max={max(l):6.1f}  {k}")
if len(sys.argv) > 3:
    for i in seg:
        if sys.argv[3] in names[i]:
            print(f"{float(rows[i]['Metric Value'].replace(',', '')):10.1f} {rows[i]['Metric Unit']} {names[i]}")
