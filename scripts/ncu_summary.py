"""Reads an ncu report (ncu -i <rep> --page raw --csv) and prints, per captured launch, the metrics DESIGN.md and
profiles/traffic.json quote: duration, DRAM bytes, L2 -> L1 bytes, tensor-pipe and DRAM utilisation, achieved occupancy,
registers, shared memory.  Usage: python scripts/ncu_summary.py gpurun_out/r2_targets.ncu-rep [--json]"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, body = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
idx = {h: i for i, h in enumerate(head)}
out = []
for r in body:
    rec = {}
    for w in want:
        if w in idx:
            rec[w] = r[idx[w]] + (" " + units[idx[w]] if units[idx[w]] else "")
    out.append(rec)
if "--json" in sys.argv:
    print(json.dumps(out, indent=1))
else:
    for rec in out:
        print("## " + rec.get("Kernel Name", "?")[:100])
        for k, v in rec.items():
            if k != "Kernel Name":
                print(f"- {k}: {v}")
        print()
