# Phase stamps of the Cholesky kernel.  Needs the measurement build of the library:
#   python -m ee274_convexcaldera_llm_quantization_b200.build --measure
#   CB_LIBRARY=libcaldera_b200_measure.so python scripts/probe_chol_phases.py
import sys; sys.path.insert(0, "/root/repo")
import torch
from ee274_convexcaldera_llm_quantization_b200 import _lib
lib=_lib.load()
for q in (224, 512):
    X=torch.randn(4096,q,device="cuda"); G0=(X.T@X).contiguous(); G=G0.clone()
    Linv=torch.empty(q,q,device="cuda"); st=torch.zeros(4,dtype=torch.int32,device="cuda")
    stamps=torch.zeros(24,dtype=torch.int64,device="cuda"); lib.cb_set_chol_timing(_lib.ptr(stamps))
    for _ in range(3):
        G.copy_(G0); lib.cb_cholesky_inverse_f32(_lib.ptr(G),q,_lib.ptr(Linv),_lib.ptr(st),_lib.stream_ptr())
    torch.cuda.synchronize(); t=stamps.tolist()
    print(f"q={q}: factor {t[1]-t[0]} clk, diag inverses {t[2]-t[1]}, triangular inverse {t[3]-t[2]}, total {t[3]-t[0]} clk")
    for pnl in range(3):
        b=4+5*pnl
        print(f"   panel {pnl}: load block {t[b+1]-t[b]}, factor block {t[b+2]-t[b+1]}, panel solve {t[b+3]-t[b+2]}, trailing update {t[b+4]-t[b+3]} clk")
    lib.cb_set_chol_timing(None)
