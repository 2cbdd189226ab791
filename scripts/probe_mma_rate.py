# Raw tcgen05.mma issue rate.  Needs the measurement build of the library:
#   python -m ee274_convexcaldera_llm_quantization_b200.build --measure
#   CB_LIBRARY=libcaldera_b200_measure.so python scripts/probe_mma_rate.py
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ee274_convexcaldera_llm_quantization_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for grid in (1, 128):
    for bn in (64, 128, 256):
        for distinct in (0, 1):
            for n in (256, 1024):
                _lib.check(lib.cb_probe_mma_rate(bn, n, distinct, grid, _lib.ptr(out), _lib.stream_ptr()), "probe")
                torch.cuda.synchronize()
                t = out.tolist()
                print(f"grid={grid} bn={bn} distinct_k={distinct} n={n}: {t[0]/n:.1f} clk/MMA total, {t[1]/n:.1f} issue; floor {bn/2:.0f}")
