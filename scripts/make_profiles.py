"""Turns the ncu outputs of scripts/gpu_profile_round.sh (gpurun_out/prof_*) into the tracked
summaries under profiles/.  Run here (no GPU needed): python scripts/make_profiles.py r1"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")

METRICS = [
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


def launches(name, title):
    src = os.path.join(G, name)
    if not os.path.exists(src):
        return
    md = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_launches.py"), src],
                        capture_output=True, text=True).stdout
    plain = os.path.join(G, name.replace("_launches.csv", "_plain.log"))
    with open(os.path.join(OUT, f"{tag}_{name.replace('prof_', '').replace('.csv', '.md')}"), "w") as f:
        f.write(f"# {title}\n\n`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` over ONE "
                "layer (scripts/profile_layer.py; per-launch times are cold-cache and serialised, compare shares).\n"
                "\"share of SM-time\" weights each launch by the fraction of the 148 SMs its grid can occupy: the\n"
                "single-CTA factorisation kernels are latency-bound and overlap with other layers in the multi-stream bench.\n\n")
        if os.path.exists(plain):
            f.write("Plain run (no profiler): `" + open(plain).read().strip() + "`\n\n")
        f.write(md)


def raw(rep, title, out_name):
    src = os.path.join(G, rep)
    if not os.path.exists(src):
        return
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, f"{tag}_{out_name}.md"), "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` ({rep}); one row per captured launch.\n\n")
        cols = [(m, lab) for m, lab in METRICS if m in hdr]
        f.write("| kernel | " + " | ".join(lab for _, lab in cols) + " |\n|---|" + "---:|" * len(cols) + "\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            vals = []
            for m, _ in cols:
                i = hdr.index(m)
                v = r[i]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                vals.append(f"{v} {units[i]}".strip())
            f.write(f"| `{name}` | " + " | ".join(vals) + " |\n")


launches("prof_layer_launches.csv", "Launch list: one 4096x4096 layer, rank 128, Q 2-bit, L/R 16-bit (BASELINE config 2), "
         "throughput execution mode (the mode bench.py runs in)")
launches("prof_layer_lr4_launches.csv", "Launch list: one 4096x4096 layer, rank 128, Q 2-bit, L/R 4-bit (LPLR loop), throughput mode")
launches("prof_layer_latency_launches.csv", "Launch list: the same layer in latency execution mode (128-CTA grids, cluster eigensolver)")
raw("prof_layer.ncu-rep", "Full captures inside the layer (tcgen05 GEMM, fused element-wise stages)", "layer_kernels")
raw("prof_quant.ncu-rep", "Full capture: quantise + pack, 4096x4096 fp32 -> 2-bit, block 64", "quant_kernel")
raw("prof_small.ncu-rep", "Full captures: Cholesky+inverse and Jacobi Rayleigh-Ritz, q = 224", "smalldense_kernels")


def traffic():
    """profiles/traffic.json: DRAM bytes per launch (read + write) of the kernels bench.py reports a roofline
    for, from the --set full captures.  bench.py copies them into `roofline.traffic`."""
    import json
    out = {}

    def rows_of(rep):
        src = os.path.join(G, rep)
        if not os.path.exists(src):
            return None, []
        txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        return rows[0], rows[2:]

    def to_bytes(v, unit):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        return float(v.replace(",", "")) * mult

    for rep, pick, key in (("prof_layer.ncu-rep", "gemm_tc_kernel", "sketch_gemm_tcgen05"),
                           ("prof_quant.ncu-rep", "quant_fast_kernel", "quantize_pack_b2_bs64")):
        src = os.path.join(G, rep)
        if not os.path.exists(src):
            continue
        txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        ki, ti = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        cand = [r for r in rows[2:] if pick in r[ki]]
        if not cand:
            continue
        r = max(cand, key=lambda r: float(r[ti].replace(",", "")))      # the sketch contraction is the longest gemm_tc launch
        out[key] = {"kernel": r[ki].split("(")[0].replace("void ", ""),
                    "dram_bytes_read": to_bytes(r[ri], units[ri]), "dram_bytes_write": to_bytes(r[wi], units[wi]),
                    "time_us_under_ncu": float(r[ti].replace(",", "")) * ({"us": 1, "ms": 1e3, "ns": 1e-3}.get(units[ti], 1)),
                    "source": f"profiles/{tag}: {rep}, ncu --set full --clock-control none"}
        out[key]["traffic"] = out[key]["dram_bytes_read"] + out[key]["dram_bytes_write"]
    with open(os.path.join(OUT, "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)


traffic()
print(os.listdir(OUT))
