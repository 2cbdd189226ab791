"""Times the L R product of the batched driver alone: batch x (4096 x 4096, K = 3 * 128 split-bf16), fp32 output through
the TMA-store epilogue of gemm_tc2 (64 MiB written per layer: the launch is bound by the store path)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ee274_convexcaldera_llm_quantization_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda"
M = N = 4096
K = 384
out = {}
for batch in (1, 8, 16, 24):
    A = torch.randn(batch, M, K, device=dev).bfloat16()
    B = torch.randn(batch, N, K, device=dev).bfloat16()
    Cs = [torch.empty(batch, M, N, device=dev) for _ in range(2)]
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    k = [0]

    def run():
        C = Cs[k[0] % 2]
        k[0] += 1
        st = lib.cb_gemm_bf16_tn_batched(batch, M, N, K, 1.0, _lib.ptr(A), K, A.stride(0) * 2, _lib.ptr(B), K, B.stride(0) * 2,
                                         _lib.ptr(C), N, C.stride(0) * 4, None, 0, 0, None, 0, 0, None, 0, None, 0, 0,
                                         _lib.ptr(counter), _lib.ptr(flag), _lib.stream_ptr())
        assert st == 0, st
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters * 1e-3
    out[f"batch{batch}"] = {"us_per_layer": t / batch * 1e6, "store_gbs": batch * M * N * 4 / t / 1e9,
                            "tflops": 2.0 * batch * M * N * K / t / 1e12}
    assert int(flag.item()) == 0
    del A, B, Cs
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
