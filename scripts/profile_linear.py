"""Times cb_packed_linear_f32 (workspace pre-allocated, CUDA events) against a dense bf16 matmul on the
reconstructed matrix, 4096 x 4096, rank 128, 2-bit codes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ee274_convexcaldera_llm_quantization_b200 import _lib

lib = _lib.load()
dev = "cuda"
m = n = 4096
r = 128
codes = torch.randint(-1, 2, (m, n), device=dev, dtype=torch.int32).to(torch.int8)
packed = torch.empty(lib.cb_packed_bytes(m * n, 2), dtype=torch.uint8, device=dev)
lib.cb_pack_codes(_lib.ptr(codes), m * n, 2, _lib.ptr(packed), _lib.stream_ptr())
L = torch.randn(m, r, device=dev)
R = torch.randn(r, n, device=dev)
s = torch.ones(1, device=dev)
W = (codes.float() + L @ R).bfloat16()
flag = torch.zeros(1, dtype=torch.int32, device=dev)
for T in (1, 16, 256, 1024, 4096):
    x = torch.randn(T, n, device=dev)
    xb = x.bfloat16()
    y = torch.empty(T, m, device=dev)
    ws = torch.empty(lib.cb_packed_linear_workspace_bytes(T, m, n, r), dtype=torch.uint8, device=dev)

    def ours():
        _lib.check(lib.cb_packed_linear_f32(_lib.ptr(x), T, n, _lib.ptr(packed), 2, _lib.ptr(s), _lib.ptr(L), _lib.ptr(R), m, r,
                                            1.0, _lib.ptr(y), _lib.ptr(flag), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "pl")

    def dense():
        return xb @ W.T
    res = []
    for fn in (ours, dense):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20 * 1e3)
    print(f"T={T}: cb_packed_linear_f32 {res[0]:.0f} us (4 MiB of codes, converts x/L/R per call), "
          f"dense bf16 matmul on 32 MiB W_hat {res[1]:.0f} us, watchdog {int(flag.item())}")
