"""Runs the quantise+pack kernel alone at the BASELINE shape (4096 x 4096 fp32 -> 2-bit, block 64) for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200 import _lib  # noqa: E402

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 2
block = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lib = _lib.load()
M = N = 4096
xs = [0.02 * torch.randn(M, N, device="cuda") for _ in range(3)]
packed = torch.empty(lib.cb_packed_bytes(M * N, bits), dtype=torch.uint8, device="cuda")
scales = torch.empty(M * N // (block or M * N), device="cuda")
for i in range(6):
    lib.cb_quantize_f32(_lib.ptr(xs[i % 3]), M, N, N, 1, bits, block, 1e-8, None, _lib.ptr(packed), _lib.ptr(scales),
                        None, _lib.stream_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    lib.cb_quantize_f32(_lib.ptr(xs[i % 3]), M, N, N, 1, bits, block, 1e-8, None, _lib.ptr(packed), _lib.ptr(scales),
                        None, _lib.stream_ptr())
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20 * 1e-3
b = 4 * M * N + M * N * bits // 8 + 4 * scales.numel()
print(f"quantize_pack bits={bits} block={block}: {t * 1e6:.1f} us, {b / t / 1e9:.0f} GB/s algorithmic")
