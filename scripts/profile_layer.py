"""One decomposition of one synthetic layer between cudaProfilerStart/Stop, for ncu:

  python scripts/profile_layer.py [--m 4096 --n 4096 --rank 128 --lbits 16 --iters 5]   # must exit 0 first
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python scripts/profile_layer.py
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from bench import synth_layer  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200 import _lib  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200.alg import make_c_params  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200.runner import CalderaLayerRunner  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=4096)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--rank", type=int, default=128)
ap.add_argument("--lbits", type=int, default=16)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--no-tc", action="store_true")
ap.add_argument("--mode", default="throughput", choices=["throughput", "latency"],
                help="library execution mode; bench.py runs in throughput mode")
a = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
_lib.set_execution_mode(a.mode)
qp = CalderaParams(Q_bits=2, L_bits=a.lbits, R_bits=a.lbits, rank=a.rank, iters=a.iters, lplr_iters=5,
                   update_order=["Q", "LR"])
cp = make_c_params(qp, True, seed=1000, use_tensor_cores=not a.no_tc)
run = CalderaLayerRunner(cp, a.m, a.n, _lib.CB_H_DIAG, dev, want_w_scaled=False)
layers = [tuple(t.to(dev) for t in synth_layer(i, a.m, a.n)) for i in range(2)]
for i in range(a.warm):
    run.enqueue(*layers[0])
torch.cuda.synchronize()
n0 = _lib.load().cb_kernel_launch_count()
torch.cuda.profiler.start()
t0 = time.perf_counter()
run.enqueue(*layers[1])
torch.cuda.synchronize()
dt = time.perf_counter() - t0
torch.cuda.profiler.stop()
small = run.read_small()
errs = small[:run.nsteps].tolist()
stats = small[run.nerr_pad + 5:run.nerr_pad + 8].view(torch.int32).tolist()
print(f"layer {a.m}x{a.n} r={a.rank} lbits={a.lbits} mode={a.mode}: {dt * 1e3:.2f} ms, "
      f"{_lib.load().cb_kernel_launch_count() - n0} launches, chol_retries={stats[0]} jacobi_sweeps={stats[1]} tc_watchdog={stats[2]}, "
      f"errors {[round(e, 5) for e in errs]}")
