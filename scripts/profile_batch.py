"""One batch of same-shape layers through the batched driver (cb_caldera_batch) between cudaProfilerStart/Stop, for ncu:

  python scripts/profile_batch.py [--m 4096 --n 4096 --rank 128 --lbits 16 --iters 5 --batch 16]   # must exit 0 first
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/batch_launches.csv python scripts/profile_batch.py
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
from ee274_convexcaldera_llm_quantization_b200.runner import BatchRunner  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=4096)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--rank", type=int, default=128)
ap.add_argument("--lbits", type=int, default=16)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
_lib.set_execution_mode("throughput")
qp = CalderaParams(Q_bits=2, L_bits=a.lbits, R_bits=a.lbits, rank=a.rank, iters=a.iters, lplr_iters=5,
                   update_order=["Q", "LR"])
cp = make_c_params(qp, True, seed=0)
run = BatchRunner(cp, a.m, a.n, _lib.CB_H_DIAG, a.batch, dev, want_packed=True)
layers = [tuple(t.to(dev) for t in synth_layer(i, a.m, a.n)) for i in range(3)]
for b in range(a.batch):                      # eager warm-up (attribute opt-ins) without a graph
    run.stage(b, *layers[b % 3])
run.enqueue()
torch.cuda.synchronize()
n0 = _lib.load().cb_kernel_launch_count()
torch.cuda.profiler.start()
t0 = time.perf_counter()
run.enqueue()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
torch.cuda.profiler.stop()
v = run.layers[0]
errs = v.small[:v.nsteps].cpu().tolist()
stats = v.small[v.nerr_pad + 4:v.nerr_pad + 8].cpu().view(torch.int32).tolist()
print(f"batch of {a.batch} layers {a.m}x{a.n} r={a.rank} lbits={a.lbits}: {dt * 1e3:.2f} ms ({dt * 1e3 / a.batch:.3f} ms per layer), "
      f"{_lib.load().cb_kernel_launch_count() - n0} launches, stats(no_iterate, chol_retries, jacobi_sweeps, watchdog)={stats}, "
      f"errors of layer 0 {[round(e, 5) for e in errs]}")
