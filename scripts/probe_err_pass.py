"""Times the error pass of the outer loop (stages.cu: err_kernel) alone at the BASELINE shape with every operand
present (Ws, int8 codes, L R product, column weights, next abs-max): 9 bytes per element.  Needs the -DCB_MEASURE build."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda"
M = N = 4096
nrot = 3
Ws = [torch.randn(M, N, device=dev) for _ in range(nrot)]
LR = [0.3 * torch.randn(M, N, device=dev) for _ in range(nrot)]
codes = [torch.randint(-1, 2, (M, N), device=dev, dtype=torch.int8) for _ in range(nrot)]
qs = torch.full((1,), 0.7, device=dev)
h = torch.rand(N, device=dev) + 0.5
num = torch.zeros(1, dtype=torch.float64, device=dev)
amax = torch.zeros(1, device=dev)


def run(i):
    st = lib.cb_probe_err_pass(_lib.ptr(Ws[i % nrot]), _lib.ptr(codes[i % nrot]), 2, _lib.ptr(qs), _lib.ptr(LR[i % nrot]),
                               _lib.ptr(h), M, N, _lib.ptr(num), _lib.ptr(amax), _lib.stream_ptr())
    assert st == 0, st


num.zero_()
run(0)
torch.cuda.synchronize()
E = Ws[0].double() - 0.7 * codes[0].double() - LR[0].double()
want = float((E * E * h.double()).sum())
print("num", float(num), "fp64 torch", want, "rel", abs(float(num) - want) / want, "amax", float(amax),
      float((Ws[0] - LR[0]).abs().max()))
for i in range(5):
    run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30):
    run(i)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 30 * 1e-3
print(f"err pass: {t * 1e6:.1f} us, {9 * M * N / t / 1e9:.0f} GB/s = {9 * M * N / t / 1e9 / 6542.1:.3f} of the copy peak")
