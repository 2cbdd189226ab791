# Phase stamps of the packed-linear consumer kernel.  Needs the measurement build of the library:
#   python -m ee274_convexcaldera_llm_quantization_b200.build --measure
#   CB_LIBRARY=libcaldera_b200_measure.so python scripts/probe_linear_parts.py
import sys; sys.path.insert(0, "/root/repo")
import torch
from ee274_convexcaldera_llm_quantization_b200 import _lib
lib=_lib.load(); dev="cuda"; m=n=4096
codes=torch.randint(-1,2,(m,n),device=dev,dtype=torch.int32).to(torch.int8)
packed=torch.empty(lib.cb_packed_bytes(m*n,2),dtype=torch.uint8,device=dev)
lib.cb_pack_codes(_lib.ptr(codes),m*n,2,_lib.ptr(packed),_lib.stream_ptr())
s=torch.ones(1,device=dev); flag=torch.zeros(1,dtype=torch.int32,device=dev)
def timeit(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
stamps=torch.zeros(12,dtype=torch.int64,device=dev)
for T in (16,4096):
    x=torch.randn(T,n,device=dev); y=torch.empty(T,m,device=dev)
    for r in (0,128):
        L=torch.randn(m,r,device=dev) if r else None; R=torch.randn(r,n,device=dev) if r else None
        ws=torch.empty(lib.cb_packed_linear_workspace_bytes(T,m,n,r),dtype=torch.uint8,device=dev)
        f=lambda: lib.cb_packed_linear_f32(_lib.ptr(x),T,n,_lib.ptr(packed),2,_lib.ptr(s),_lib.ptr(L) if r else None,_lib.ptr(R) if r else None,m,r,1.0,_lib.ptr(y),_lib.ptr(flag),_lib.ptr(ws),ws.numel(),_lib.stream_ptr())
        print(f"T={T} r={r}: {timeit(f):.0f} us")
        lib.cb_set_gemm_timing(_lib.ptr(stamps)); f(); torch.cuda.synchronize(); lib.cb_set_gemm_timing(None)
        t=stamps.tolist(); e=t[0]
        print(f"    main kernel CTA 0: prologue {t[1]-e}, first stage complete {t[3]-e}, expansion loop done {t[2]-e}, "
              f"last stage complete {t[4]-e}, accumulators complete {t[5]-e}, epilogue done {t[6]-e}, exit {t[7]-e} clk")
    xb=torch.empty(T,n,device=dev,dtype=torch.bfloat16)
    print(f"T={T} to_bf16(x) alone: {timeit(lambda: lib.cb_convert_bf16(_lib.ptr(x),T,n,n,_lib.ptr(xb),n,None,0,None,_lib.stream_ptr())):.0f} us")
