"""The two kernels bench.py reports a roofline for, launched a few times at the workload's shape for one
`ncu --set full --clock-control none --import-source on` capture (run plain first; it must exit 0):

  python scripts/ncu_targets.py
  ncu --set full --clock-control none --import-source on -k regex:'gemm_tc2_kernel|quant_stream_kernel|quant_fast_kernel' -c 6 \
      -o gpurun_out/r2_targets python scripts/ncu_targets.py

Operands are rotated so that no launch finds its input in L2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ee274_convexcaldera_llm_quantization_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda"
M, N, q, nb = 4096, 4096, 224, int(os.environ.get("CB_NCU_BATCH", "23"))
As = [(0.02 * torch.randn(nb, M, N, device=dev)).bfloat16() for _ in range(3)]
B = torch.randn(nb, q, N, device=dev).bfloat16()
Cb = torch.empty(nb, M, q, device=dev, dtype=torch.bfloat16)
Ct = torch.empty(nb, q, M, device=dev, dtype=torch.bfloat16)
counter = torch.zeros(2, dtype=torch.int32, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
for A in As:
    st = lib.cb_gemm_bf16_tn_batched(nb, M, q, N, 1.0, _lib.ptr(A), N, A.stride(0) * 2, _lib.ptr(B), N, B.stride(0) * 2,
                                     None, 0, 0, _lib.ptr(Cb), q, Cb.stride(0) * 2, _lib.ptr(Ct), M, Ct.stride(0) * 2,
                                     None, 0, None, 0, 0, _lib.ptr(counter), _lib.ptr(flag), _lib.stream_ptr())
    assert st == 0
torch.cuda.synchronize()
assert int(flag.item()) == 0
del As, B, Cb, Ct
xs = [0.02 * torch.randn(M, N, device=dev) for _ in range(3)]
packed = torch.empty(M * N // 4, dtype=torch.uint8, device=dev)
scales = torch.empty(M * N // 64, device=dev)
for x in xs:
    lib.cb_quantize_f32(_lib.ptr(x), M, N, N, 1, 2, 64, 1e-8, None, _lib.ptr(packed), _lib.ptr(scales), None,
                        _lib.stream_ptr())
torch.cuda.synchronize()
print("ok")
