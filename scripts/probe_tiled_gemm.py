"""Sketch contraction Zt[q, m] = Pt[q, n] * Y[m, n]^T with Y row-major vs tiled (64 x 64 tiles), Y HBM-cold."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ee274_convexcaldera_llm_quantization_b200 import _lib

lib = _lib.load()
dev = "cuda"
M = N = 4096
q = 224
ys = [torch.randn(M, N, device=dev).bfloat16() for _ in range(6)]
yt = [y.view(M // 64, 64, N // 64, 64).permute(0, 2, 1, 3).contiguous() for y in ys]
Pt = torch.randn(q, N, device=dev).bfloat16()
Zt = torch.empty(q, M, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)


def run(mats, layout, iters=30):
    k = 0
    for _ in range(3):
        lib.cb_gemm_bf16_tn(q, M, N, 1.0, _lib.ptr(Pt), N, _lib.ptr(mats[k % len(mats)]), N, _lib.ptr(Zt), M, 1, layout,
                            _lib.ptr(flag), None, 0, _lib.stream_ptr()); k += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.cb_gemm_bf16_tn(q, M, N, 1.0, _lib.ptr(Pt), N, _lib.ptr(mats[k % len(mats)]), N, _lib.ptr(Zt), M, 1, layout,
                            _lib.ptr(flag), None, 0, _lib.stream_ptr()); k += 1
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for cl in (1, 2, 4, 8):
    lib.cb_set_gemm_cluster(cl)
    ctas = 120
    print(f"cluster={cl}: row-major cold {run(ys, 0):.1f} us  warm {run(ys[:1], 0):.1f} us | tiled cold {run(yt, 2):.1f} us  warm {run(yt[:1], 2):.1f} us")
a = Zt.clone()
lib.cb_gemm_bf16_tn(q, M, N, 1.0, _lib.ptr(Pt), N, _lib.ptr(ys[0]), N, _lib.ptr(Zt), M, 1, 0, _lib.ptr(flag), None, 0, _lib.stream_ptr())
b = Zt.clone()
lib.cb_gemm_bf16_tn(q, M, N, 1.0, _lib.ptr(Pt), N, _lib.ptr(yt[0]), N, _lib.ptr(Zt), M, 1, 2, _lib.ptr(flag), None, 0, _lib.stream_ptr())
torch.cuda.synchronize()
print("equal:", torch.equal(b, Zt), "flag", int(flag.item()))
lib.cb_set_gemm_cluster(1)
lib.cb_gemm_bf16_tn(q, M, N, 1.0, _lib.ptr(Pt), N, _lib.ptr(ys[0]), N, _lib.ptr(Zt), M, 1, 0, _lib.ptr(flag), None, 0, _lib.stream_ptr())
torch.cuda.synchronize()
print("cluster 8 == cluster 1:", torch.equal(b, Zt))
