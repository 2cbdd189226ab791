"""profiles/r2_sass_summary.md: static counts of the tcgen05 / TMA / TMEM / bulk-copy instructions per kernel of the shipped
library.  python scripts/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ee274_convexcaldera_llm_quantization_b200", "libcaldera_b200.so")
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "UTCBAR", "UTCATOMSWS"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
cur, hmma = None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Za-z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    if op.startswith("HMMA"):
        hmma += 1
    for k in KEYS:
        if op.startswith(k):
            per.setdefault(cur, collections.OrderedDict()).setdefault(k, collections.Counter())[op] += 1
names = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
print("# Round 2: tcgen05 / TMA / TMEM / bulk-copy instructions in the shipped library (SASS)\n")
print("`python scripts/sass_summary.py` = `cuobjdump -sass ee274_convexcaldera_llm_quantization_b200/libcaldera_b200.so` (sm_100a,")
print("release build), static instruction counts per kernel; only kernels that contain at least one of the mnemonics are")
print("listed (the recipe in B200_PROFILING.md: `UTCHMMA` = tcgen05.mma, `UTMALDG` / `UTMASTG` = cp.async.bulk.tensor load /")
print("store, `UBLKCP` = cp.async.bulk (1-D bulk copy global -> shared), `LDTM` = tcgen05.ld, `UTCBAR` = tcgen05.commit,")
print("`UTCATOMSWS` = tcgen05.alloc / dealloc; `.2CTA` = cta_group::2).\n")
print("| kernel | " + " | ".join(KEYS) + " |")
print("|---|" + "---|" * len(KEYS))
seen = set()
for mangled, name in zip(per, names):
    name = re.sub(r"\((?:int|bool)\)", "", name)
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("cb::", "").strip()
    row = []
    for k in KEYS:
        c = per[mangled].get(k)
        row.append(", ".join(f"{op} x{n}" for op, n in c.items()) if c else "-")
    line = f"| `{short}` | " + " | ".join(row) + " |"
    if line not in seen:
        seen.add(line)
        print(line)
print(f"\n`HMMA` (mma.sync) occurrences in the whole library: {hmma}.")
