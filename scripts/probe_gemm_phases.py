"""Phase breakdown (clock64 stamps of CTA 0) of the sketch contraction for tile widths / cluster sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ee274_convexcaldera_llm_quantization_b200 import _lib
lib = _lib.load()
dev = "cuda"
M = N = 4096
q = 224
Y = torch.randn(M, N, device=dev).bfloat16()
Pt = torch.randn(q, N, device=dev).bfloat16()
Zt = torch.empty(q, M, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
stamps = torch.zeros(12, dtype=torch.int64, device=dev)
lib.cb_set_gemm_timing(_lib.ptr(stamps))
names = ["prologue", "->last load issued", "first stage landed (from entry)", "last stage landed (from entry)",
         "accumulator complete (from entry)", "epilogue done (from entry)", "exit (from entry)"]
for ctas, kbs, split, layout, tag in ((120, 1, 1, 0, "bn64 kb1"), (120, 2, 1, 0, "bn64 kb2"), (120, 2, 1, 1, "bn64 kb2 loads-only"),
                                      (64, 1, 1, 0, "bn128 kb1"), (64, 2, 1, 0, "bn128 kb2"), (64, 2, 2, 0, "bn128 kb2 split2"),
                                      (32, 1, 1, 0, "bn256"), (32, 1, 4, 0, "bn256 split4")):
    lib.cb_set_gemm_target_ctas(ctas); lib.cb_set_gemm_kblocks(kbs)
    for _ in range(3):
        lib.cb_gemm_bf16_tn(q, M, N, 1.0, _lib.ptr(Pt), N, _lib.ptr(Y), N, _lib.ptr(Zt), M, split, layout,
                            _lib.ptr(flag), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    torch.cuda.synchronize()
    t = stamps.tolist()
    e = t[0]
    print(f"{tag}: prologue {t[1]-e}, last-load-issued {t[2]-e}, first-landed {t[3]-e}, last-landed {t[4]-e}, "
          f"acc-complete {t[5]-e}, epilogue-done {t[6]-e}, exit {t[7]-e} clk")
# the same contraction with the layer's bf16 epilogues (row-major + transposed), staged vs direct stores
Zb = torch.empty(q, M, device=dev, dtype=torch.bfloat16)
Zbt = torch.empty(M, q, device=dev, dtype=torch.bfloat16)
for ctas, staged, tag in ((120, 1, "bn64 kb2 bf16-out staged"), (120, 0, "bn64 kb2 bf16-out direct"),
                          (32, 1, "bn256 bf16-out staged"), (32, 0, "bn256 bf16-out direct")):
    lib.cb_set_gemm_target_ctas(ctas); lib.cb_set_gemm_kblocks(2); lib.cb_set_gemm_staged_epilogue(staged)
    for _ in range(3):
        lib.cb_gemm_bf16_tn_bf16out(q, M, N, 1.0, _lib.ptr(Pt), N, _lib.ptr(Y), N, _lib.ptr(Zb), M, _lib.ptr(Zbt), q,
                                    None, None, _lib.ptr(flag), _lib.stream_ptr())
    torch.cuda.synchronize()
    t = stamps.tolist()
    e = t[0]
    print(f"{tag}: prologue {t[1]-e}, first-landed {t[3]-e}, last-landed {t[4]-e}, acc-complete {t[5]-e}, "
          f"staged-in-smem {t[8]-e}, barrier {t[9]-e}, epilogue-done {t[6]-e}, exit {t[7]-e} clk")
lib.cb_set_gemm_staged_epilogue(1)
lib.cb_set_gemm_timing(None)
