#!/usr/bin/env python
"""bench.py -- decomposition throughput of the CALDERA hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): decomposed matrices / s at 4096 x 4096, rank 128, 2-bit Q
(config[1]: L/R 16-bit, activation aware, 5 outer iterations, update_order Q,LR); second half of the metric:
wall seconds of the full Llama-2-7B-shape job (config[3]), emitted as `full_7b_wall_s` on the same line.
One step = `--slots` x `--batch` independent layers (one full caldera() decomposition each): `--slots` CUDA-graph
replays in flight, each advancing `--batch` same-shape layers in lock step (cb_caldera_batch).

  value     matrices/s with inputs resident in HBM (CUDA-event timed, max over ranks)
  e2e       the same through the public API (caldera_async) with HOST (pinned) inputs: H2D of W and h, the
            decomposition, D2H of the packed result, all inside the timed region; ONE host thread per rank
  full_7b_wall_s   224 layers layer-sharded over the ranks + gather of the packed blobs on rank 0, for L/R 16-bit and
            4-bit (inputs generated on the owning GPU before the clock, as SURVEY 8e prescribes)
  parity    exact-match fraction of the Q codes (iterate 0 and best iterate) and relative difference of the best
            error against the stored run of the unmodified reference on the same layer (tests/golden/fullsize_c2*)
  roofline  the dominant kernel, timed alone with CUDA events at the workload's shape
  cpu_baseline  the reference on the host cores, bounded sample
  reference_cuda  informational: the unmodified reference with device="cuda" (torch / cuBLAS / cuSOLVER) on this GPU

`--impl reference` times the UNMODIFIED reference's caldera(device="cpu") (oracle/_ref, copied there by
__graft_entry__.build() in the build container; the numpy port of oracle/ if that copy is absent) and prints the
same line with "impl": "reference".
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
REF_DIR = os.path.join(ROOT, "oracle", "_ref")      # the unmodified reference sources (git-ignored, see oracle/README)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def pin_host_threads(cores):
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use all host threads."""
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = str(cores)
    import torch
    torch.set_num_threads(cores)


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, "src", "caldera", "decomposition", "alg.py"))


def import_reference():
    """The reference's own modules from oracle/_ref (its `src` package name collides with this repo's drop-in
    shim, so it is imported under a private name)."""
    import importlib.util
    import types
    if "caldera_ref" in sys.modules:
        return sys.modules["caldera_ref"]
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_DIR)
    try:
        alg = importlib.import_module("src.caldera.decomposition.alg")
        dc = importlib.import_module("src.caldera.utils.dataclasses")
        qz = importlib.import_module("src.caldera.utils.quantization")
    finally:
        sys.path.remove(REF_DIR)
        ref_mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
    ns = types.SimpleNamespace(caldera=alg.caldera, CalderaParams=dc.CalderaParams, QuantizerFactory=qz.QuantizerFactory,
                               modules=ref_mods)
    sys.modules["caldera_ref"] = ns
    return ns


def reference_layer(device, m=M, n=N, rank=RANK, iters=ITERS, idx=0):
    """One full caldera() call of the UNMODIFIED reference on the workload's layer.  Returns (seconds, result)."""
    import torch
    ref = import_reference()
    W, h = synth_layer(idx, m, n)
    qf = ref.QuantizerFactory(method="uniform", block_size=64)
    p = ref.CalderaParams(compute_quantized_component=True, compute_low_rank_factors=True, Q_bits=2, L_bits=16, R_bits=16,
                          rank=rank, iters=iters, lplr_iters=5, activation_aware_LR=True, update_order=["Q", "LR"],
                          quant_factory_Q=qf, quant_factory_LR=qf, rand_svd=False, sigma_reg=0.0)
    H = torch.diag(h)                       # the caller contract of main.py:165
    torch.manual_seed(42)
    if device != "cpu":
        W, H = W.to(device), H.to(device)
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = ref.caldera(p, W, H, device=device, use_tqdm=False, scale_W=True)
    if device != "cpu":
        torch.cuda.synchronize()
    return time.perf_counter() - t0, d


def port_sample():
    """Fallback when oracle/_ref is absent: one outer iteration of the numpy port, scaled linearly to ITERS."""
    import numpy as np
    from oracle import caldera_oracle as orc
    W, h = synth_layer(0)
    p = orc.OracleParams(Q_bits=2, L_bits=16, R_bits=16, rank=RANK, iters=1, update_order=["Q", "LR"])
    t0 = time.perf_counter()
    orc.caldera_oracle(p, W.numpy(), np.diag(h.numpy()))
    return (time.perf_counter() - t0) * ITERS


def cpu_sample():
    """(seconds per layer, kind, description) of one bounded CPU sample of the workload."""
    if reference_available():
        dt, d = reference_layer("cpu")
        return dt, "reference", (f"one full caldera(device='cpu') call of the unmodified reference on a {M}x{N} layer "
                                 f"(all {ITERS} iterations, exact SVD, 4-product error): {dt:.1f} s; best error "
                                 f"{min(d.errors['LR']):.6f}")
    dt = port_sample()
    return dt, "port", (f"oracle/_ref absent: 1 outer iteration of the numpy port on a {M}x{N} layer scaled x{ITERS}: {dt:.1f} s")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    pin_host_threads(cores)
    budget = 200.0
    t_start = time.perf_counter()
    if args.warmup > 0 and reference_available():
        reference_layer("cpu", 256, 256, 16, 1)         # pages the libraries in; a full layer is ~1 minute
    times, kind, desc = [], "port", ""
    for _ in range(max(args.steps, 1)):
        dt, kind, desc = cpu_sample()
        times.append(dt)
        if time.perf_counter() - t_start + dt > budget:
            break
    per_layer = sum(times) / len(times)
    value = 1.0 / per_layer
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_layer * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "steps_measured": len(times),
                       "step": "one layer (the GPU arm's step is a batch of layers; the metric is per matrix either way)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{len(times)} sample(s); last: {desc}"},
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


def roofline_probes(dev, peaks, nbatch):
    """Dominant kernel classes timed alone (CUDA events on the launching stream) at the workload's shapes; operands
    are rotated so that consecutive launches never find their input in the 126 MB L2."""
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

    # (a') the same kernel on a 4x larger tensor (8192 x 8192): at 4096 x 4096 a launch lasts ~17 us, of which the ramp
    #      up and the tail are a fixed ~3-4 us; this entry shows the rate the kernel streams at once that is amortised
    big = [0.02 * torch.randn(2 * M, 2 * N, device=dev) for _ in range(2)]
    packed_b = torch.empty(numel, dtype=torch.uint8, device=dev)
    scales_b = torch.empty(numel // 16, device=dev)

    def quant_big():
        x = big[k[0] % 2]
        k[0] += 1
        lib.cb_quantize_f32(_lib.ptr(x), 2 * M, 2 * N, 2 * N, 1, 2, 64, 1e-8, None, _lib.ptr(packed_b), _lib.ptr(scales_b), None,
                            _lib.stream_ptr())
    t = time_kernel(quant_big)
    out["quantize_pack_b2_bs64_8192x8192"] = {"bound": "hbm", "achieved": 4 * bytes_q / t / 1e9, "peak": peaks["hbm_gbs"],
                                              "unit": "GB/s", "frac": 4 * bytes_q / t / 1e9 / peaks["hbm_gbs"], "traffic": None,
                                              "seconds": t, "algorithmic_bytes": 4 * bytes_q}
    del big, packed_b, scales_b

    # (b) whole-tensor quantise + pack (the form caldera() itself uses, alg.py:247): two passes
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
    del xs
    # (c) the sketch contraction Z[m, q] = Y[m, K = n] * P[q, K = n]^T of the rank-r step as the batched driver
    #     launches it: ONE launch of the CTA-pair tcgen05 kernel for all layers of a batch (bf16 operands in HBM,
    #     bf16 outputs in both orientations) -- the kernel with the largest share of GPU time (profiles/).
    q = 224
    skinny = ("skinny: 2q flop per 2-byte element of Y = 224 flop/B, so min(tensor peak, AI x HBM) = "
              "min(1662.7, 224 x 6.54) = 1465 TFLOP/s; 256 x 224 tiles on CTA pairs (cta_group::2), persistent, two TMEM "
              "accumulators")
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    for label, nb in (("sketch_gemm_tcgen05", nbatch), ("sketch_gemm_tcgen05_batch32", 32), ("sketch_gemm_tcgen05_batch1", 1)):
        nrot = 2 if nb >= 8 else 8            # >= 2 x nb x 32 MiB of A operands: never L2 resident across launches
        As = [(0.02 * torch.randn(nb, M, N, device=dev)).bfloat16() for _ in range(nrot)]
        B = torch.randn(nb, q, N, device=dev).bfloat16()
        Cb = torch.empty(nb, M, q, device=dev, dtype=torch.bfloat16)
        Ct = torch.empty(nb, q, M, device=dev, dtype=torch.bfloat16)

        def sketch():
            A = As[k[0] % nrot]
            k[0] += 1
            lib.cb_gemm_bf16_tn_batched(nb, M, q, N, 1.0, _lib.ptr(A), N, A.stride(0) * 2, _lib.ptr(B), N, B.stride(0) * 2,
                                        None, 0, 0, _lib.ptr(Cb), q, Cb.stride(0) * 2, _lib.ptr(Ct), M, Ct.stride(0) * 2,
                                        None, 0, None, 0, 0, _lib.ptr(counter), _lib.ptr(flag), _lib.stream_ptr())
        t = time_kernel(sketch, iters=10)
        flops = 2.0 * nb * M * N * q
        out[label] = {"bound": "tensor", "achieved": flops / t / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                      "frac": flops / t / 1e12 / peaks["bf16_tflops"], "traffic": None, "seconds": t,
                      "algorithmic_flops": flops, "layers_per_launch": nb,
                      "hbm_gbs": nb * (2 * M * N + 2 * q * N + 2 * 2 * q * M) / t / 1e9, "note": skinny}
        del As, B, Cb, Ct
        torch.cuda.empty_cache()
    assert int(flag.item()) == 0
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
    from ee274_convexcaldera_llm_quantization_b200 import _lib, parity
    from ee274_convexcaldera_llm_quantization_b200 import model_job as mj
    from ee274_convexcaldera_llm_quantization_b200 import scheduler as sch
    from ee274_convexcaldera_llm_quantization_b200.alg import make_c_params, caldera, caldera_async
    from ee274_convexcaldera_llm_quantization_b200.engine import get_engine, release_engines
    import ctypes as C
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

    fac = QuantizerFactory(method="uniform", block_size=64)

    def params_for(lbits):
        return CalderaParams(Q_bits=2, L_bits=lbits, R_bits=lbits, rank=RANK, iters=ITERS, lplr_iters=5,
                             activation_aware_LR=True, update_order=["Q", "LR"], quant_factory_Q=fac,
                             quant_factory_LR=fac, rand_svd=False, sigma_reg=0)
    qp = params_for(16)
    cp = make_c_params(qp, True, seed=0)

    # synthetic layers: distinct per rank (weak scaling: per-GPU work is fixed), pinned on the host
    npool = 3
    host_layers = []
    for i in range(npool):
        W, h = synth_layer(rank * npool + i)
        host_layers.append((W.pin_memory(), h.pin_memory()))
    dev_layers = [(W.to(dev), h.to(dev)) for W, h in host_layers]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident timing (value): `slots` graph replays in flight, each advancing `batch` layers in lock step
    # (cb_caldera_batch); inputs resident in HBM, staged into the batch's slabs by device-side copies; one host thread
    nslots, nbatch = max(1, args.slots), max(1, args.batch)
    nstreams = nslots * nbatch                          # layers in flight
    streams = [torch.cuda.Stream(device=dev) for _ in range(nslots)]
    from ee274_convexcaldera_llm_quantization_b200.runner import BatchRunner
    assert lib.cb_caldera_batch_supported(C.byref(cp), M, N, _lib.CB_H_DIAG), "workload must take the batched driver"
    runners = [BatchRunner(cp, M, N, _lib.CB_H_DIAG, nbatch, dev, want_packed=True) for _ in range(nslots)]
    for s_, r_ in zip(streams, runners):
        with torch.cuda.stream(s_):
            r_.capture()
    torch.cuda.synchronize()

    def run_resident(first, count):
        """`count` batches (of nbatch layers) round robin over the slots."""
        main_stream = torch.cuda.current_stream()
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record(main_stream)
        for s_ in streams:
            s_.wait_event(start)
        for i in range(count):
            r_ = runners[i % nslots]
            with torch.cuda.stream(streams[i % nslots]):
                for b in range(nbatch):
                    r_.stage(b, *dev_layers[(first + i * nbatch + b) % npool])
                r_.replay([1000 + rank] * nbatch)
        for s_ in streams:
            ev = torch.cuda.Event()
            ev.record(s_)
            main_stream.wait_event(ev)
        stop.record(main_stream)
        return start, stop

    batch = nstreams                      # one step = `nslots` batches of `nbatch` independent layers
    run_resident(0, args.warmup * nslots)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = lib.cb_kernel_launch_count()
    e0, e1 = run_resident(args.warmup * nslots, args.steps * nslots)
    barrier()
    launches = lib.cb_kernel_launch_count() - launches0
    secs = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    v0 = runners[0].layers[0]
    errs = v0.small[:v0.nsteps].cpu().tolist()
    # latency of one batch alone, and of a lone layer (batch of one), for reference
    torch.cuda.synchronize()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    runners[0].replay([1000 + rank] * nbatch)
    l1.record()
    torch.cuda.synchronize()
    batch_latency_ms = l0.elapsed_time(l1)
    kernels_per_layer = runners[0].graph_kernels / nbatch
    del runners, streams, v0
    torch.cuda.empty_cache()
    lone = BatchRunner(cp, M, N, _lib.CB_H_DIAG, 1, dev, want_packed=True)
    lone.capture()
    lone.stage(0, *dev_layers[0])
    torch.cuda.synchronize()
    l0.record()
    lone.replay([1000 + rank])
    l1.record()
    torch.cuda.synchronize()
    layer_latency_ms = l0.elapsed_time(l1)
    del lone
    torch.cuda.empty_cache()

    # ---- end-to-end timing through the public API: ONE host thread, pinned host buffers in, packed result out.
    # caldera_async() stages W and h with asynchronous H2D copies on the layer's stream, replays the layer's graph,
    # and the `consume` hook enqueues the D2H copies of the packed codes and the factors into pinned host buffers;
    # the ~100-byte result record (error trajectory, scales) follows.  `.result()` is called a batch later, so the
    # host never waits for the layer it has just submitted.
    # the engine's arenas: slots x batch slabs of ~0.72 GiB (99 GiB at 6 x 23); the library's default cap is 60 % of
    # the free memory, the bench owns the whole GPU
    os.environ.setdefault("CB_ENGINE_MAX_BYTES", str(int(0.85 * torch.cuda.mem_get_info(dev)[0])))
    engine = get_engine(dev, nslots, nbatch)
    out_hosts = [{"Q_packed": torch.empty(M * N // 4, dtype=torch.uint8).pin_memory(),
                  "L": torch.empty(M, RANK).pin_memory(), "R": torch.empty(RANK, N).pin_memory()}
                 for _ in range(nstreams)]

    def submit_e2e(i):
        W, h = host_layers[i % npool]
        dst = out_hosts[i % nstreams]

        def to_host(run, kept):
            dst["Q_packed"].copy_(run.Q_packed, non_blocking=True)
            dst["L"].copy_(run.L, non_blocking=True)
            dst["R"].copy_(run.R, non_blocking=True)
        return caldera_async(qp, W, h, device=dev, use_tqdm=False, W_copy="none", seed=1000 + rank, return_dense=False,
                             return_packed=False, consume=to_host, slots=nslots, batch=nbatch)

    e2e_marks = []

    def run_e2e(first, count):
        pending, last = [], None
        del e2e_marks[:]
        for i in range(count):
            pending.append(submit_e2e(first + i))
            if len(pending) > nstreams:                  # harvest a layer submitted a whole batch ago
                last = pending.pop(0).result()
            if (i + 1) % nstreams == 0:
                e2e_marks.append(time.perf_counter())    # submit-side clock, one mark per step
        engine.flush()
        for hd in pending:
            last = hd.result()
        return last
    # context for the e2e number: what the host link delivers to this GPU for one pinned 256 MiB buffer
    probe_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    probe_d.copy_(probe_h, non_blocking=True)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(4):
        probe_d.copy_(probe_h, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = 4 * probe_h.numel() / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del probe_h, probe_d
    e2e_steps = max(1, args.steps)
    run_e2e(0, max(1, min(args.warmup, 2)) * batch)       # every slot captures its graph
    barrier()
    t0 = time.perf_counter()
    last_dec = run_e2e(args.warmup * batch, e2e_steps * batch)
    torch.cuda.synchronize()
    e2e_secs = time.perf_counter() - t0
    e2e_step_secs = [round(b - a, 4) for a, b in zip([t0] + e2e_marks[:-1], e2e_marks)]
    barrier()
    e2e_errors = {k: [round(e, 6) for e in v] for k, v in last_dec.errors.items()}
    del last_dec                                  # its factors are views of a slot's arena

    tmax = torch.tensor([secs, e2e_secs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    secs, e2e_secs = (float(x) for x in tmax.tolist())
    h2d = batch * (M * N * 4 + N * 4)
    d2h = batch * (M * N // 4 + (M + N) * RANK * 4 + 4 * (ITERS * 2 + 20))

    # ---- parity against the stored run of the unmodified reference on this very layer (rank 0, layer seed 1000)
    parity_rep = None
    if rank == 0 and not args.no_parity:
        try:
            z, _ = parity.load_fullsize_golden("c2")
            W0, h0 = host_layers[0]
            qp0 = params_for(16)
            qp0.iters, qp0.update_order = 1, ["Q"]            # the first Q update alone: iterate 0
            d0 = caldera(qp0, W0, h0, device=dev, use_tqdm=False, W_copy="none", global_scale=z["global_scale"],
                         use_cuda_graph=True)
            dref = caldera(qp, W0, h0, device=dev, use_tqdm=False, W_copy="none", global_scale=z["global_scale"],
                           use_cuda_graph=True, seed=1000)
            parity_rep = parity.parity_report(dref, "c2", d0)
            parity_rep["best_step"], parity_rep["best_step_reference"] = dref.best_step, z["best_step"]
            del d0, dref
        except FileNotFoundError:
            parity_rep = {"unavailable": "tests/golden/fullsize_c2.json not found"}

    # ---- second half of the metric: the full Llama-2-7B-shape job (config 4), L/R 16-bit and 4-bit
    full7b = None
    if not args.no_model:
        import gc
        release_engines()
        del out_hosts, engine
        gc.collect()
        torch.cuda.empty_cache()
        sys.stderr.write(f"[bench] before the model-level job: {torch.cuda.memory_allocated(dev) / 2**30:.1f} GiB allocated, "
                         f"{torch.cuda.mem_get_info(dev)[0] / 2**30:.1f} GiB free\n")
        names, shapes = mj.llama_shapes(args.model_blocks)
        full7b = {"layers": len(names), "params": sum(m * n for m, n in shapes), "rank": RANK, "iters": ITERS,
                  "streams": args.model_streams, "slots": args.model_slots, "inputs": "generated on the owning GPU before the clock (SURVEY 8e)",
                  "timed": "barrier -> all layers decomposed -> packed blobs gathered on rank 0 -> device synchronised; "
                           "max over ranks; median of five passes after a first one that captures the CUDA graphs"}
        shards = sch.shard_layout(params_for(16), shapes, world)[0]
        store = mj.synth_layers(shapes, shards[rank], dev)
        sch.warm_up_gather(dev, dst=0)
        for lbits in (16, 4):
            prm = params_for(lbits)
            res, passes, mallocs = None, [], []
            for attempt in range(6):          # pass 0 captures the CUDA graphs; passes 1-5 are timed, the median is reported
                res = None
                ms0 = torch.cuda.memory_stats(dev)
                res = mj.run_model_job(prm, names, shapes, store, rank, world, dev, streams=args.model_streams,
                                       slots=args.model_slots, barrier=barrier)
                ms1 = torch.cuda.memory_stats(dev)
                tt = torch.tensor([res["decompose_s"], res["gather_s"], res["wall_s"]], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                passes.append([float(x) for x in tt.tolist()] + [int(res["graphs_captured"])])
                # cudaMalloc / cudaFree calls of torch's allocator during the pass (a cudaFree synchronises the device)
                mallocs.append([int(ms1.get(k_, 0) - ms0.get(k_, 0)) for k_ in ("num_device_alloc", "num_device_free", "num_alloc_retries")])
            first_pass = passes[0][2]
            med = sorted(passes[1:], key=lambda t_: t_[2])[len(passes[1:]) // 2]
            key = f"lr{lbits}"
            full7b[key] = {"wall_s": med[2], "decompose_s": med[0], "gather_s": med[1],
                           "timed_passes_wall_s": [round(t_[2], 4) for t_ in passes[1:]],
                           "best_pass_wall_s": min(t_[2] for t_ in passes[1:]),
                           "allocator_device_alloc_free_retries_per_pass_rank0": mallocs,
                           "gathered_bytes": int(res["gathered_bytes"]), "first_pass_wall_s_incl_graph_capture": first_pass,
                           "graphs_captured_in_timed_pass": max(t_[3] for t_ in passes[1:])}
            if rank == 0:
                parts = sch.split_gathered(res["arena"], res["shards"], res["sizes"])
                first = sch.unpack_decomposition(parts[0])
                full7b[key]["layers_gathered"] = len(parts)
                full7b[key]["first_layer_best_error"] = min(first["errors"]["LR"])
            res = None
            release_engines()
            torch.cuda.empty_cache()
        del store
        torch.cuda.empty_cache()

    if rank == 0:
        value = world * args.steps * batch / secs
        e2e_value = world * e2e_steps * batch / e2e_secs
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 I/O, bf16 tensor-core operands, f32 accumulate (f32 small factorisations)", "data": "synthetic",
                "config": {"workload": WORKLOAD, "l2": "inputs larger than L2: every layer in flight has a ~0.5 GiB working set (126 MB L2), 3 input "
                                 "layers rotated; the roofline probes rotate >= 192 MB of operands",
                           "layers_per_step": batch, "layers_in_flight": nstreams, "graph_replays_in_flight": nslots,
                           "layers_per_replay": nbatch, "e2e_host_threads": 1, "batch_latency_ms": batch_latency_ms,
                           "e2e_result": "packed Q codes + scale, L, R, error trajectory",
                           "hw_queues": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "execution_mode": args.mode, "cuda_graphs": True, "single_layer_latency_ms": layer_latency_ms,
                           "kernels_per_layer": kernels_per_layer,
                           "parallelism": f"layer-sharded x{world}, no data-path collective",
                           "sketch_width": 224, "power_iters": "12 cold + 3 per warm-started step", "peaks": peaks["source"]},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "submit_side_step_seconds": e2e_step_secs, "host_link_h2d_gbs_measured": h2d_gbs,
                        "h2d_gbs_used": e2e_value / world * (M * N * 4 + N * 4) / 1e9,
                        "note": "fp32 W crosses the host link once per layer (67 MB): e2e per GPU is bounded by "
                                "host_link_h2d_gbs_measured / 0.0671 matrices/s"},
                "gpu_launches": int(launches) * world, "gpu_launches_per_rank": int(launches), "clocks": clocks,
                "errors_last_layer": [round(e, 6) for e in errs],
                "e2e_errors_last_layer": e2e_errors,
                "parity": parity_rep}
        if full7b is not None:
            line["full_7b_wall_s"] = {k: v["wall_s"] for k, v in full7b.items() if isinstance(v, dict)}
            line["full_7b"] = full7b
        if True:                                   # rank 0, every N: the dominant kernel timed alone on this GPU
            probes = roofline_probes(dev, peaks, nbatch)
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
            if world == 1 and not args.no_cpu:
                cores = host_threads()
                pin_host_threads(cores)
                dt, kind, desc = cpu_sample()
                line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
            if world == 1 and not args.no_ref_cuda and reference_available():
                # informational: what a user of the reference gets today on this GPU (torch / cuBLAS / cuSOLVER)
                try:
                    reference_layer(str(dev), 512, 512, 16, 1)
                    dt, dref = reference_layer(str(dev))
                    line["reference_cuda"] = {"value": 1.0 / dt, "unit": UNIT, "seconds_per_layer": dt,
                                              "best_error": float(min(dref.errors["LR"])),
                                              "what": "the unmodified reference's caldera(device='cuda') on the same layer, one call"}
                except Exception as exc:      # noqa: BLE001  (informational leg must not take the bench down)
                    line["reference_cuda"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
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
    ap.add_argument("--no-model", action="store_true", help="skip the full Llama-2-7B-shape job (full_7b_wall_s)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity report against the stored reference run")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the informational reference-on-CUDA leg")
    ap.add_argument("--model-blocks", type=int, default=32, help="transformer blocks of the model-level job (32 = Llama-2-7B)")
    ap.add_argument("--model-streams", type=int, default=64, help="layers in flight per GPU in the model-level job")
    ap.add_argument("--slots", type=int, default=6, help="graph replays in flight per GPU")
    ap.add_argument("--model-slots", type=int, default=4, help="graph replays in flight per GPU in the model-level job")
    ap.add_argument("--batch", type=int, default=23,
                    help="same-shape layers advancing in lock step per graph replay (23 x 16 tiles of the sketch contraction "
                         "= 4.97 waves of the 74 CTA pairs; 24 would be 5.19, i.e. a sixth, mostly empty wave)")
    ap.add_argument("--mode", default="throughput", choices=["throughput", "latency"],
                    help="execution mode written into cb_caldera_params.exec_mode (single-layer driver only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
