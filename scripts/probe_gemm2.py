"""Times the batched CTA-pair contraction (cb_gemm_bf16_tn_batched) at the sketch shape of the headline workload:
batch x (4096 x 224 x 4096), bf16 outputs in both orientations, operands rotated so that they do not sit in L2."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ee274_convexcaldera_llm_quantization_b200 import _lib

lib = _lib.load()
dev = "cuda"
M, N, K = 4096, 224, 4096
out = {}
BATCHES = tuple(int(a) for a in sys.argv[1:]) or (1, 4, 16, 32, 37)
for batch in BATCHES:
    nrot = max(2, 256 // batch // 8 + 1) if batch < 16 else 2
    As = [torch.randn(batch, M, K, device=dev).bfloat16() for _ in range(nrot)]
    B = torch.randn(batch, N, K, device=dev).bfloat16()
    Cb = torch.empty(batch, M, N, device=dev, dtype=torch.bfloat16)
    Ct = torch.empty(batch, N, M, device=dev, dtype=torch.bfloat16)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    k = [0]
    def run(mc=0):
        A = As[k[0] % nrot]; k[0] += 1
        st = lib.cb_gemm_bf16_tn_batched(batch, M, N, K, 1.0, _lib.ptr(A), K, A.stride(0) * 2, _lib.ptr(B), K, B.stride(0) * 2,
                                         None, 0, 0, _lib.ptr(Cb), N, Cb.stride(0) * 2, _lib.ptr(Ct), M, Ct.stride(0) * 2,
                                         None, 0, None, 0, mc, _lib.ptr(counter), _lib.ptr(flag), _lib.stream_ptr())
        assert st == 0, st
    for mc in ((0,) if batch != 32 else (0, 64, 32)):
        for _ in range(3): run(mc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters): run(mc)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / iters * 1e-3
        flops = 2.0 * batch * M * N * K
        out[f"batch{batch}_clusters{mc}"] = {"seconds": t, "tflops": flops / t / 1e12, "per_layer_us": t / batch * 1e6}
    assert int(flag.item()) == 0
    del As, B, Cb, Ct
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
