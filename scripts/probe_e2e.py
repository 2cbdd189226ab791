"""Where the end-to-end path (pinned host W -> caldera_async -> packed result in pinned host memory) spends its time:
per-step wall times over several steps, how long the submitting thread was blocked waiting for a slot, and the same
loop with the input copies taken out (W already on the device) or the output copies taken out.

  python scripts/probe_e2e.py --slots 4 --batch 24 --steps 6
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_layer, M, N, RANK, ITERS  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200 import _lib  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200.alg import caldera_async  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200.engine import get_engine, release_engines  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams  # noqa: E402
from src.caldera.utils.quantization import QuantizerFactory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=3)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--modes", default="full,no_h2d,no_d2h")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
_lib.set_execution_mode("throughput")
fac = QuantizerFactory(method="uniform", block_size=64)
qp = CalderaParams(Q_bits=2, L_bits=16, R_bits=16, rank=RANK, iters=ITERS, lplr_iters=5, activation_aware_LR=True,
                   update_order=["Q", "LR"], quant_factory_Q=fac, quant_factory_LR=fac, rand_svd=False, sigma_reg=0)
host_layers = [tuple(t.pin_memory() for t in synth_layer(i)) for i in range(3)]
dev_layers = [(W.to(dev), h.to(dev)) for W, h in host_layers]
nstreams = a.slots * a.batch
out_hosts = [{"Q_packed": torch.empty(M * N // 4, dtype=torch.uint8).pin_memory(),
              "L": torch.empty(M, RANK).pin_memory(), "R": torch.empty(RANK, N).pin_memory()} for _ in range(nstreams)]
engine = get_engine(dev, a.slots, a.batch)


def run(mode, count):
    pending, blocked, nblocked, marks = [], 0.0, 0, []
    t_begin = time.perf_counter()
    for i in range(count):
        W, h = (dev_layers if mode == "no_h2d" else host_layers)[i % 3]
        dst = out_hosts[i % nstreams]

        def to_host(run_, kept, dst=dst):
            if mode != "no_d2h":
                dst["Q_packed"].copy_(run_.Q_packed, non_blocking=True)
                dst["L"].copy_(run_.L, non_blocking=True)
                dst["R"].copy_(run_.R, non_blocking=True)
        t0 = time.perf_counter()
        pending.append(caldera_async(qp, W, h, device=dev, use_tqdm=False, W_copy="none", seed=1000, return_dense=False,
                                     return_packed=False, consume=to_host, slots=a.slots, batch=a.batch))
        if len(pending) > nstreams:
            pending.pop(0).result()
        dt = time.perf_counter() - t0
        if dt > 2e-3:
            blocked += dt
            nblocked += 1
        if (i + 1) % nstreams == 0:
            marks.append(time.perf_counter() - t_begin)
    engine.flush()
    for hd in pending:
        hd.result()
    torch.cuda.synchronize()
    total = time.perf_counter() - t_begin
    steps = [marks[0]] + [marks[k] - marks[k - 1] for k in range(1, len(marks))]
    return total, blocked, nblocked, steps


if __name__ == "__main__":
    for mode in a.modes.split(","):
        run(mode, 2 * nstreams)
        total, blocked, nblocked, steps = run(mode, a.steps * nstreams)
        print(f"{a.slots}x{a.batch} {mode:7s}: {a.steps * nstreams / total:7.1f} matrices/s, total {total:.3f} s, submitting thread blocked "
              f"{blocked:.3f} s in {nblocked} calls, submit-side step times {[round(s, 3) for s in steps]}", flush=True)
    release_engines()
