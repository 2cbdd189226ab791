"""Runs the Cholesky+inverse and the Jacobi eigensolver kernels alone (q x q), for ncu."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200 import _lib  # noqa: E402

q = int(sys.argv[1]) if len(sys.argv) > 1 else 224
lib = _lib.load()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(q)
A = torch.randn(q, 4096, generator=g, device=dev, dtype=torch.float64)
A = A * torch.linspace(1.0, 0.6, q, device=dev, dtype=torch.float64)[:, None]
G0 = (A @ A.T / 4096).float().contiguous()
Linv = torch.empty(q, q, device=dev)
status = torch.zeros(2, dtype=torch.int32, device=dev)
evals = torch.empty(q, device=dev)
evecs = torch.empty(q, q, device=dev)
work = torch.empty(q * q + q + 8, device=dev)
for rep in range(3):
    G = G0.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lib.cb_cholesky_inverse_f32(_lib.ptr(G), q, _lib.ptr(Linv), _lib.ptr(status), _lib.stream_ptr())
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    lib.cb_jacobi_eigh_from_chol_f32(_lib.ptr(G), q, _lib.ptr(evals), _lib.ptr(evecs), _lib.ptr(work),
                                     _lib.ptr(status[1:]), _lib.stream_ptr())
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"q={q} rep={rep}: chol+inv {1e6 * (t1 - t0):.1f} us, jacobi {1e6 * (t2 - t1):.1f} us, "
          f"retries={int(status[0])} sweeps={int(status[1])}")
ref = torch.linalg.eigvalsh(G0.double()).flip(0)
print("eval rel err", float((evals.double() - ref).abs().max() / ref.max()))
