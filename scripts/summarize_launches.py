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
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
tot = 0.0
tot_sm = 0.0
NUM_SMS = 148


def grid_ctas(text):
    nums = [int(x) for x in re.findall(r"\d+", text or "1")]
    out = 1
    for v in nums:
        out *= v
    return max(out, 1)
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").strip()
    a = agg[name]
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
    sm = v * min(grid_ctas(row.get("Grid Size")), NUM_SMS) / NUM_SMS   # time x fraction of the SMs a launch can occupy
    a[3] += sm
    tot += v
    tot_sm += sm
print("| kernel | launches | total ms | avg us | max us | share of serial time | share of SM-time |")
print("|---|---:|---:|---:|---:|---:|---:|")
for k, (c, t, mx, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {t / 1e3:.3f} | {t / c:.1f} | {mx:.1f} | {100 * t / tot:.1f}% | {100 * sm / tot_sm:.1f}% |")
print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot / 1e3:.3f} | | | 100% | 100% (= {tot_sm / 1e3:.3f} ms of a full GPU) |")
