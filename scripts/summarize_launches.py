"""Per-kernel summary (markdown) of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/xxx_launches.md
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").strip()
    a = agg[name]
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
    tot += v
print(f"| kernel | launches | total ms | avg us | max us | share |")
print("|---|---:|---:|---:|---:|---:|")
for k, (c, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {t / 1e3:.3f} | {t / c:.1f} | {mx:.1f} | {100 * t / tot:.1f}% |")
print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot / 1e3:.3f} | | | 100% |")
