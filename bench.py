#!/usr/bin/env python
"""bench.py -- decomposition throughput of the CALDERA hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): decomposed matrices / s at 4096 x 4096, rank 128, 2-bit Q
(config[1]: L/R 16-bit, activation aware, 5 outer iterations, update_order Q,LR).
One step = one batch of `--streams` independent layers (one full caldera() decomposition each).

  value     matrices/s with inputs resident in HBM (cb_caldera_layer enqueued back to back,
            CUDA-event timed, max over ranks)
  e2e       the same through the public API with HOST (pinned) inputs: H2D of W and h, the
            decomposition, D2H of the packed result, all inside the timed region
  roofline  the dominant kernel, timed alone with CUDA events at the workload's shape
  cpu_baseline  the numpy port of the reference (oracle/) on the host cores, bounded sample

`--impl reference` times the oracle port only (the Python reference cannot travel to the
GPU box) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# Independent layers run on their own CUDA streams.  The driver multiplexes streams onto 8 hardware
# work queues by default, which falsely serialises layers once more than 8 are in flight; 32 is the
# maximum.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M = N = 4096
RANK = 128
ITERS = 5
WORKLOAD = "llama2-7b q_proj 4096x4096, rank 128, Q 2-bit, L/R 16-bit, activation-aware, iters=5, order Q,LR"
METRIC = "decomp matrices/sec (4096x4096, rank-128, 2-bit Q)"
UNIT = "matrices/s"


def synth_layer(idx, m=M, n=N):
    """SURVEY.md section 8d synthetic inputs: per-layer seeded W ~ 0.02 N(0,1), h = 0.5 + U(0,1)."""
    import torch
    g = torch.Generator().manual_seed(1000 + idx)
    W = 0.02 * torch.randn(m, n, generator=g, dtype=torch.float32)
    h = 0.5 + torch.rand(n, generator=g, dtype=torch.float32)
    return W, h


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU arm
def cpu_reference_sample(threads=None):
    """One outer iteration (Q update + LR update, exact SVD, four-product error evaluation)
    of the workload's layer with the numpy port of the reference; scaled linearly to ITERS."""
    import numpy as np
    from oracle import caldera_oracle as orc
    W, h = synth_layer(0)
    p = orc.OracleParams(Q_bits=2, L_bits=16, R_bits=16, rank=RANK, iters=1, update_order=["Q", "LR"])
    t0 = time.perf_counter()
    d = orc.caldera_oracle(p, W.numpy(), np.diag(h.numpy()))
    dt = time.perf_counter() - t0
    return dt, d


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    times = []
    budget = 240.0
    t_start = time.perf_counter()
    if args.warmup > 0:
        cpu_reference_sample()      # one warm-up sample is enough for a ~10 s CPU step
    for _ in range(max(args.steps, 1)):
        dt, _ = cpu_reference_sample()
        times.append(dt)
        if time.perf_counter() - t_start + dt > budget:
            break
    per_iter = sum(times) / len(times)
    value = 1.0 / (per_iter * ITERS)
    sample = (f"{len(times)} sample(s) of 1 outer iteration (Q+LR, exact SVD) of one {M}x{N} layer, "
              f"{per_iter:.2f} s each, scaled x{ITERS} iterations")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_iter * ITERS * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "steps_measured": len(times)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- GPU arm
def time_kernel(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def roofline_probes(dev, peaks):
    """Dominant kernel classes timed alone (CUDA events on the launching stream) at the
    workload's shapes.  Inputs (64 MiB W, 192 MiB rotating) exceed nothing smaller than L2,
    so three tensors are rotated to defeat the 126 MB L2."""
    import torch
    from ee274_convexcaldera_llm_quantization_b200 import _lib
    lib = _lib.load()
    out = {}
    xs = [0.02 * torch.randn(M, N, device=dev) for _ in range(3)]
    numel = M * N
    # (a) quantise + pack, 2-bit, block 64 (the direct LowMemoryQuantizer path)
    packed = torch.empty(numel // 4, dtype=torch.uint8, device=dev)
    scales = torch.empty(numel // 64, device=dev)
    k = [0]

    def quant():
        x = xs[k[0] % 3]
        k[0] += 1
        lib.cb_quantize_f32(_lib.ptr(x), M, N, N, 1, 2, 64, 1e-8, None, _lib.ptr(packed), _lib.ptr(scales), None,
                            _lib.stream_ptr())
    t = time_kernel(quant)
    bytes_q = 4 * numel + numel // 4 + 4 * numel // 64
    out["quantize_pack_b2_bs64"] = {"bound": "hbm", "achieved": bytes_q / t / 1e9, "peak": peaks["hbm_gbs"],
                                    "unit": "GB/s", "frac": bytes_q / t / 1e9 / peaks["hbm_gbs"], "traffic": None,
                                    "seconds": t, "algorithmic_bytes": bytes_q}
    # (b) sketch contraction Zt[q, m] = Pt[q, K=n] * Y[m, K=n]^T on the tcgen05 kernel: the kernel with
    #     the largest share of GPU time per layer (ncu launch list, profiles/).  bf16 operands in HBM.
    q = 224
    # six bf16 operands (6 x 32 MiB = 192 MiB) so that the rotation does not fit in the 126 MB L2
    ys = [x.bfloat16() for x in xs] + [(0.02 * torch.randn(M, N, device=dev)).bfloat16() for _ in range(3)]
    Pt = torch.randn(q, N, device=dev).bfloat16()
    Zt = torch.empty(q, M, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)

    # outputs as in the layer: bf16 row-major (q x m) and transposed (m x q), no fp32 copy
    Zts = [torch.empty(q, M, device=dev, dtype=torch.bfloat16) for _ in range(4)]
    Zs = [torch.empty(M, q, device=dev, dtype=torch.bfloat16) for _ in range(4)]
    flops = 2.0 * M * N * q

    def sketch(j=0):
        y = ys[k[0] % 6]
        k[0] += 1
        lib.cb_gemm_bf16_tn_bf16out(q, M, N, 1.0, _lib.ptr(Pt), N, _lib.ptr(y), N, _lib.ptr(Zts[j]), M, _lib.ptr(Zs[j]), q,
                                    None, None, _lib.ptr(flag), _lib.stream_ptr())

    def entry(t, note, **extra):
        e = {"bound": "tensor", "achieved": flops / t / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
             "frac": flops / t / 1e12 / peaks["bf16_tflops"], "traffic": None, "seconds": t,
             "algorithmic_flops": flops, "hbm_gbs": (2 * M * N + 2 * q * N + 2 * 2 * q * M) / t / 1e9, "note": note}
        e.update(extra)
        return e
    skinny = ("skinny: 2q flop per 2-byte element of Y = 224 flop/B, so min(tensor peak, AI x HBM) = "
              "min(1662.7, 224 x 6.54) = 1465 TFLOP/s; N = 256 tiles run the tensor pipe at its rate, the rest is "
              "prologue + epilogue of a 64-K-block tile")
    # as it runs in the timed region (throughput mode): 32-CTA grids of 128 x 256 tiles, four of them
    # side by side on different streams; seconds = elapsed / launches
    lib.cb_set_gemm_target_ctas(32)
    side = [torch.cuda.Stream(device=dev) for _ in range(4)]
    for j, s_ in enumerate(side):
        with torch.cuda.stream(s_):
            for _ in range(3):
                sketch(j)
    torch.cuda.synchronize()
    rounds = 20
    # fork-join graph: 4 streams x `rounds` launches, so that the host's launch rate is not what is timed
    cap = torch.cuda.Stream(device=dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
        for j, s_ in enumerate(side):
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                for _ in range(rounds):
                    sketch(j)
        for s_ in side:
            torch.cuda.current_stream().wait_stream(s_)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    t4 = e0.elapsed_time(e1) * 1e-3 / (rounds * 4)
    out["sketch_gemm_tcgen05"] = entry(t4, skinny + "; 32-CTA grid, 4 launches in flight (as in the timed region)",
                                       grid_ctas=32, in_flight=4)
    out["sketch_gemm_tcgen05_32cta_alone"] = entry(time_kernel(sketch, iters=20), skinny + "; 32-CTA grid alone "
                                                   "(occupies 32 of 148 SMs)", grid_ctas=32, in_flight=1)
    lib.cb_set_gemm_target_ctas(120)
    out["sketch_gemm_tcgen05_128cta_alone"] = entry(time_kernel(sketch, iters=20), skinny + "; latency-mode grid "
                                                    "(128 x 64 tiles, 128 CTAs) alone", grid_ctas=128, in_flight=1)
    _lib.set_execution_mode(_lib.execution_mode())     # restore the grid policy of the current mode
    assert int(flag.item()) == 0
    # (c) whole-tensor quantise + pack (the form caldera() itself uses, alg.py:247): two passes
    def quant_whole():
        x = xs[k[0] % 3]
        k[0] += 1
        lib.cb_quantize_f32(_lib.ptr(x), M, N, N, 1, 2, 0, 1e-8, None, _lib.ptr(packed), _lib.ptr(scales), None,
                            _lib.stream_ptr())
    t = time_kernel(quant_whole)
    bytes_w = 8 * numel + numel // 4
    out["quantize_pack_b2_whole"] = {"bound": "hbm", "achieved": bytes_w / t / 1e9, "peak": peaks["hbm_gbs"],
                                     "unit": "GB/s", "frac": bytes_w / t / 1e9 / peaks["hbm_gbs"], "traffic": None,
                                     "seconds": t, "algorithmic_bytes": bytes_w}
    return out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ee274_convexcaldera_llm_quantization_b200 import _lib
    from ee274_convexcaldera_llm_quantization_b200.alg import make_c_params, caldera
    from ee274_convexcaldera_llm_quantization_b200.runner import CalderaLayerRunner
    from src.caldera.utils.dataclasses import CalderaParams
    from src.caldera.utils.quantization import QuantizerFactory

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    peaks = load_peaks()
    # many independent layers in flight: the library's throughput mode (small contraction grids that
    # overlap across streams, single-CTA eigensolver); "latency" is the single-layer optimum
    _lib.set_execution_mode(args.mode)
    if args.gemm_ctas > 0:
        lib.cb_set_gemm_target_ctas(args.gemm_ctas)

    fac = QuantizerFactory(method="uniform", block_size=64)
    qp = CalderaParams(Q_bits=2, L_bits=16, R_bits=16, rank=RANK, iters=ITERS, lplr_iters=5,
                       activation_aware_LR=True, update_order=["Q", "LR"], quant_factory_Q=fac,
                       quant_factory_LR=fac, rand_svd=False, sigma_reg=0)
    cp = make_c_params(qp, True, seed=1000 + rank)

    # synthetic layers: distinct per rank (weak scaling: per-GPU work is fixed), pinned on the host
    npool = 3
    host_layers = []
    for i in range(npool):
        W, h = synth_layer(rank * npool + i)
        host_layers.append((W.pin_memory(), h.pin_memory()))
    dev_layers = [(W.to(dev), h.to(dev)) for W, h in host_layers]
    # One runner (outputs + ~0.5 GiB workspace) per stream.  Layers are independent, so S of them
    # are kept in flight: the single-CTA factorisation kernels of one layer (Cholesky, Jacobi)
    # overlap with the bandwidth/tensor-bound kernels of the others.
    nstreams = max(1, args.streams)
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    runners = [CalderaLayerRunner(cp, M, N, _lib.CB_H_DIAG, dev, want_packed=True, want_w_scaled=False)
               for _ in range(nstreams)]
    runner = runners[0]
    if not args.no_graph:
        for s_, r_ in zip(streams, runners):
            with torch.cuda.stream(s_):
                r_.capture()
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_resident(first, count):
        main = torch.cuda.current_stream()
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record(main)
        for s_ in streams:
            s_.wait_event(start)
        for i in range(count):
            W, h = dev_layers[(first + i) % npool]
            with torch.cuda.stream(streams[i % nstreams]):
                if args.no_graph:
                    runners[i % nstreams].enqueue(W, h)
                else:
                    runners[i % nstreams].launch(W, h, seed=1000 + rank)
        for s_ in streams:
            ev = torch.cuda.Event()
            ev.record(s_)
            main.wait_event(ev)
        stop.record(main)
        return start, stop

    # ---- resident timing (value)
    run_resident(0, args.warmup * nstreams)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = lib.cb_kernel_launch_count()
    # one step = one batch of `nstreams` independent layers (one per stream)
    batch = nstreams
    e0, e1 = run_resident(args.warmup * batch, args.steps * batch)
    barrier()
    launches = lib.cb_kernel_launch_count() - launches0
    secs = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    errs = runner.read_small()[:runner.nsteps].tolist()
    # single-stream latency of one layer, for reference
    torch.cuda.synchronize()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    if args.no_graph:
        runner.enqueue(*dev_layers[0])
    else:
        runner.launch(*dev_layers[0], seed=1000 + rank)
    l1.record()
    torch.cuda.synchronize()
    layer_latency_ms = l0.elapsed_time(l1)

    # ---- end-to-end timing through the public API, host buffers in, packed result out
    import concurrent.futures as cf
    auto_workers = min(24, max(4, 3 * host_threads() // (2 * max(world, 1))))
    nworkers = max(1, min(nstreams, args.e2e_workers if args.e2e_workers > 0 else auto_workers))
    out_hosts = [{"Q_packed": torch.empty(M * N // 4, dtype=torch.uint8).pin_memory(),
                  "L": torch.empty(M, RANK).pin_memory(), "R": torch.empty(RANK, N).pin_memory()}
                 for _ in range(nworkers)]
    e2e_streams = [torch.cuda.Stream(device=dev) for _ in range(nworkers)]

    def step_e2e(i):
        w = i % nworkers
        W, h = host_layers[i % npool]
        torch.cuda.set_device(dev)
        with torch.cuda.stream(e2e_streams[w]):
            d = caldera(qp, W, h, device=dev, use_tqdm=False, W_copy="none", seed=1000 + rank,
                        use_cuda_graph=not args.no_graph, return_dense=not args.e2e_packed_only)
            out_hosts[w]["Q_packed"].copy_(d.Q_packed, non_blocking=True)
            out_hosts[w]["L"].copy_(d.L, non_blocking=True)
            out_hosts[w]["R"].copy_(d.R, non_blocking=True)
            e2e_streams[w].synchronize()
        return d.errors["LR"][-1]

    def run_e2e(first, count):
        # worker w handles steps w, w + nworkers, ... sequentially on its own stream
        def work(w):
            return [step_e2e(first + i) for i in range(w, count, nworkers)]
        with cf.ThreadPoolExecutor(max_workers=nworkers) as ex:
            return list(ex.map(work, range(nworkers)))
    e2e_steps = max(1, args.steps)
    run_e2e(0, max(1, min(args.warmup, 2)) * batch)   # every worker warms its stream, workspace and graph
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.warmup * batch, e2e_steps * batch)
    torch.cuda.synchronize()
    e2e_secs = time.perf_counter() - t0
    barrier()

    tmax = torch.tensor([secs, e2e_secs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    secs, e2e_secs = (float(x) for x in tmax.tolist())
    h2d = batch * (M * N * 4 + N * 4)
    d2h = batch * (M * N // 4 + (M + N) * RANK * 4 + runner.small.numel() * 4)

    if rank == 0:
        value = world * args.steps * batch / secs
        e2e_value = world * e2e_steps * batch / e2e_secs
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "l2": "inputs larger than L2: every layer in flight has a ~0.5 GiB working set (126 MB L2), 3 input "
                                 "layers rotated; the roofline probes rotate >= 192 MB of operands",
                           "layers_per_step": batch, "layers_in_flight": nstreams, "e2e_host_threads": nworkers, "e2e_result": "packed Q codes + scale, L, R" if args.e2e_packed_only else "dense + packed",
                           "hw_queues": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "execution_mode": args.mode, "cuda_graphs": not args.no_graph, "single_layer_latency_ms": layer_latency_ms,
                           "parallelism": f"layer-sharded x{world}, no data-path collective",
                           "sketch_width": 224, "power_iters": "12 cold + 3 per warm-started step", "peaks": peaks["source"]},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps},
                "gpu_launches": int(launches) * world, "gpu_launches_per_rank": int(launches), "clocks": clocks,
                "errors_last_layer": [round(e, 6) for e in errs]}
        if world == 1:
            probes = roofline_probes(dev, peaks)
            dominant = os.environ.get("CB_DOMINANT", "sketch_gemm_tcgen05")
            tpath = os.path.join(ROOT, "profiles", "traffic.json")      # ncu --set full, DRAM bytes per launch
            if os.path.exists(tpath):
                with open(tpath) as f:
                    for name, rec in json.load(f).items():
                        if name in probes:
                            probes[name]["traffic"] = rec["traffic"]
            line["roofline"] = {k: probes[dominant][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
            line["roofline"]["kernel"] = dominant
            line["roofline_all"] = probes
            if not args.no_cpu:
                cores = host_threads()
                dt, _ = cpu_reference_sample()
                v = 1.0 / (dt * ITERS)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"1 outer iteration (Q+LR, exact SVD, 4-product error) of one "
                                                  f"{M}x{N} layer with the numpy port: {dt:.2f} s, scaled x{ITERS}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of replaying CUDA graphs")
    ap.add_argument("--e2e-dense", dest="e2e_packed_only", action="store_false",
                    help="e2e leg: also materialise the dense fp32 Q / int8 codes copies in the returned decomposition "
                         "(default: packed codes + factors only, which is what is copied back to the host)")
    ap.add_argument("--mode", default="throughput", choices=["throughput", "latency"],
                    help="library execution mode (cb_set_execution_mode)")
    ap.add_argument("--gemm-ctas", type=int, default=0, help="grid-size target of the tcgen05 contractions (0 = library default)")
    ap.add_argument("--streams", type=int, default=32, help="independent layers kept in flight per GPU")
    ap.add_argument("--e2e-workers", type=int, default=0,
                    help="host threads driving the public API in the e2e leg (0 = min(24, 1.5 x host cores / ranks), at least 4)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
