"""Do single-CTA factorisation kernels from different streams overlap?  S streams x R launches each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ee274_convexcaldera_llm_quantization_b200 import _lib
lib = _lib.load()
dev = "cuda"
q = 224
R = 20
for kind in ("chol", "jacobi"):
    for S in (1, 2, 4, 8, 16):
        streams = [torch.cuda.Stream() for _ in range(S)]
        X = [torch.randn(4096, q, device=dev) for _ in range(S)]
        G0 = [(x.T @ x).contiguous() for x in X]
        G = [g.clone() for g in G0]
        Linv = [torch.empty(q, q, device=dev) for _ in range(S)]
        status = [torch.zeros(4, dtype=torch.int32, device=dev) for _ in range(S)]
        ev = [torch.empty(q, device=dev) for _ in range(S)]
        V = [torch.empty(q, q, device=dev) for _ in range(S)]
        work = [torch.empty(q * q + q + 8, device=dev) for _ in range(S)]
        # factor once so that jacobi has a valid Cholesky factor as input
        for i in range(S):
            lib.cb_cholesky_inverse_f32(_lib.ptr(G[i]), q, _lib.ptr(Linv[i]), _lib.ptr(status[i]), _lib.stream_ptr())
        torch.cuda.synchronize()
        def body(i):
            if kind == "chol":
                G[i].copy_(G0[i])
                lib.cb_cholesky_inverse_f32(_lib.ptr(G[i]), q, _lib.ptr(Linv[i]), _lib.ptr(status[i]), _lib.stream_ptr())
            else:
                lib.cb_jacobi_eigh_from_chol_f32(_lib.ptr(G[i]), q, _lib.ptr(ev[i]), _lib.ptr(V[i]), _lib.ptr(work[i]),
                                                 _lib.ptr(status[i]), _lib.stream_ptr())
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                body(i)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for r in range(R if kind == "chol" else 4):
            for i, s in enumerate(streams):
                with torch.cuda.stream(s):
                    body(i)
        for s in streams:
            e = torch.cuda.Event(); e.record(s); torch.cuda.current_stream().wait_event(e)
        e1.record(); torch.cuda.synchronize()
        n = R if kind == "chol" else 4
        print(f"{kind}: {S} streams x {n} launches: {e0.elapsed_time(e1):.2f} ms  ({e0.elapsed_time(e1) / n * 1e3:.0f} us per round)")
