"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > k.csv
    python scripts/ncu_lines.py k.csv [top]
"""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
samples = collections.Counter()
insts = collections.Counter()
text = {}
hdr = None
fname = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ns, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= max(ns, ie):
        continue
    try:
        key = (fname, int(r[0]))
        samples[key] += int(r[ns])
        insts[key] += int(r[ie])
        text[key] = r[1]
    except ValueError:
        pass
tot = sum(samples.values()) or 1
print("total samples", tot, "total warp-instructions", sum(insts.values()))
for key, s in samples.most_common(top):
    print(f"{s:7d} {100 * s / tot:5.1f}%  inst={insts[key]:9d}  {key[0]}:{key[1]:<4d} {text[key].strip()[:100]}")
