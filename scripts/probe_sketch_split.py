"""Sketch contraction Zt[q, m] = Pt[q, n] * Y[m, n]^T: N-tile width x K split sweep (two-launch split-K)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ee274_convexcaldera_llm_quantization_b200 import _lib

lib = _lib.load()
dev = "cuda"
M = N = 4096
q = int(sys.argv[1]) if len(sys.argv) > 1 else 224
ys = [torch.randn(M, N, device=dev).bfloat16() for _ in range(4)]
Pt = torch.randn(q, N, device=dev).bfloat16()
Zt = torch.empty(q, M, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)


def run(splitk, iters=30, layout=0):
    k = 0
    def one():
        nonlocal k
        lib.cb_gemm_bf16_tn(q, M, N, 1.0, _lib.ptr(Pt), N, _lib.ptr(ys[k % 4]), N, _lib.ptr(Zt), M, splitk, layout,
                            _lib.ptr(flag), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()); k += 1
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        one()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for ctas, name in ((120, "bn64"), (64, "bn128"), (32, "bn256")):
  for kbs in (1, 2):
    lib.cb_set_gemm_target_ctas(ctas); lib.cb_set_gemm_kblocks(kbs)
    print(name, f"kb{kbs}", " ".join(f"split{s}: {run(s):.1f} us" for s in (1, 2, 4, 8)))
for ctas, name in ((120, "bn64"), (64, "bn128"), (32, "bn256")):
    lib.cb_set_gemm_target_ctas(ctas)
    print(name, "loads only:", " ".join(f"split{s}: {run(s, layout=1):.1f} us" for s in (1, 2, 4)))
print("flag", int(flag.item()))
