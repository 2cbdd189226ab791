"""Sweeps the geometry of quant_stream_kernel (warps x ring stages, with / without programmatic stream serialisation)
against the register-staged quant_fast_kernel at the BASELINE shape, and checks that both write identical bytes.

Needs the -DCB_MEASURE build (CB_LIBRARY=.../libcaldera_b200_measure.so): only that build reads CB_QS_*.
CB_QS_WARPS=0 selects the register-staged kernel."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda"
PEAK = 6542.1


def run(x, bits, block, packed, scales, codes=None, deq=None):
    st = lib.cb_quantize_f32(_lib.ptr(x), x.shape[0], x.shape[1], x.shape[1], 1, bits, block, 1e-8,
                             None if codes is None else _lib.ptr(codes), _lib.ptr(packed), _lib.ptr(scales),
                             None if deq is None else _lib.ptr(deq), _lib.stream_ptr())
    assert st == 0, st


def setcfg(warps, stages, pdl):
    os.environ["CB_QS_WARPS"], os.environ["CB_QS_STAGES"], os.environ["CB_QS_PDL"] = str(warps), str(stages), str(pdl)


out = {}
for (M, N) in ((4096, 4096), (8192, 8192)):
    xs = [0.02 * torch.randn(M, N, device=dev) for _ in range(3)]
    xs[0][0, :64] = 0.0                      # an all-zero block (scale = eps)
    xs[0][1, 0] = 1e-30
    numel = M * N
    for bits, block in ((2, 64), (4, 64), (2, 0), (8, 128), (2, 256), (4, 32)):
        if (M, N) != (4096, 4096) and (bits, block) != (2, 64):
            continue
        nsc = numel // (block or numel)
        ref_p = torch.empty(lib.cb_packed_bytes(numel, bits), dtype=torch.uint8, device=dev)
        ref_s = torch.empty(nsc, device=dev)
        ref_c = torch.empty(numel, dtype=torch.int8, device=dev)
        ref_d = torch.empty(numel, device=dev)
        setcfg(0, 3, 0)
        run(xs[0], bits, block, ref_p, ref_s, ref_c, ref_d)
        for cfg in ((0, 3, 0), (16, 2, 1), (16, 2, 0), (24, 2, 1), (20, 2, 1), (24, 1, 1), (12, 3, 1)):
            setcfg(*cfg)
            p, s = torch.zeros_like(ref_p), torch.zeros_like(ref_s)
            c, d = torch.zeros_like(ref_c), torch.zeros_like(ref_d)
            run(xs[0], bits, block, p, s, c, d)
            torch.cuda.synchronize()
            same = bool(torch.equal(p, ref_p) and torch.equal(s, ref_s) and torch.equal(c, ref_c) and torch.equal(d, ref_d))
            for i in range(5):
                run(xs[i % 3], bits, block, p, s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 30
            e0.record()
            for i in range(iters):
                run(xs[i % 3], bits, block, p, s)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / iters * 1e-3
            by = 4 * numel + numel * bits // 8 + 4 * nsc + (4 * numel if block == 0 else 0)
            key = f"{M}x{N}_b{bits}_bs{block}_w{cfg[0]}_s{cfg[1]}_pdl{cfg[2]}"
            out[key] = {"us": round(t * 1e6, 2), "gbs": round(by / t / 1e9, 1), "frac": round(by / t / 1e9 / PEAK, 3), "identical": same}
            print(key, out[key], flush=True)
    del xs
    torch.cuda.empty_cache()
bad = [k for k, v in out.items() if not v["identical"]]
print(json.dumps({"mismatching": bad}))
