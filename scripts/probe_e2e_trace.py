"""GPU-side timeline of the end-to-end path, per batch: how long the input copies of a batch took on its slot's
stream, how long the batch then computed (graph replay + output copies), and how long the submitting thread spent in
each caldera_async() call.  Finds out whether a slow e2e step is the host link, the GPU or the host thread.

  python scripts/probe_e2e_trace.py --slots 6 --batch 16 --steps 8
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synth_layer, M, N, RANK, ITERS  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200 import _lib, engine as eng_mod  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200.alg import caldera_async  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200.engine import get_engine, release_engines  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams  # noqa: E402
from src.caldera.utils.quantization import QuantizerFactory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=6)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
_lib.set_execution_mode("throughput")
fac = QuantizerFactory(method="uniform", block_size=64)
qp = CalderaParams(Q_bits=2, L_bits=16, R_bits=16, rank=RANK, iters=ITERS, lplr_iters=5, activation_aware_LR=True,
                   update_order=["Q", "LR"], quant_factory_Q=fac, quant_factory_LR=fac, rand_svd=False, sigma_reg=0)
host_layers = [tuple(t.pin_memory() for t in synth_layer(i)) for i in range(3)]
nstreams = a.slots * a.batch
out_hosts = [{"Q_packed": torch.empty(M * N // 4, dtype=torch.uint8).pin_memory(),
              "L": torch.empty(M, RANK).pin_memory(), "R": torch.empty(RANK, N).pin_memory()} for _ in range(nstreams)]
engine = get_engine(dev, a.slots, a.batch)
trace = []          # per batch: dict of events
origin = torch.cuda.Event(enable_timing=True)
issue_log = []      # (slot, t_begin, t_replayed, t_end) of every launcher-thread issue

orig_issue = eng_mod.LayerEngine._issue
orig_replay = eng_mod.BatchRunner.replay


def traced_replay(self, seeds):
    t0 = time.perf_counter()
    orig_replay(self, seeds)
    self._t_replay = (t0, time.perf_counter())


def traced_issue(self, g):
    rec = getattr(g, "_trace", None)
    t0 = time.perf_counter()
    if rec is not None:
        rec["staged"] = torch.cuda.Event(enable_timing=True)
        rec["staged"].record(g.slot.stream)
        rec["t_launch"] = t0
    runner = g.runner
    orig_issue(self, g)
    t1 = time.perf_counter()
    tr = getattr(runner, "_t_replay", (t0, t0))
    issue_log.append((g.slot.index, t0, tr[0], tr[1], t1))
    if rec is not None:
        rec["t_launched"] = t1
        rec["done"] = torch.cuda.Event(enable_timing=True)
        rec["done"].record(g.slot.stream)


eng_mod.LayerEngine._issue = traced_issue
eng_mod.BatchRunner.replay = traced_replay


def run(count, keep):
    pending, calls = [], []
    if os.environ.get("PROBE_NO_GC") == "1":
        import gc
        gc.disable()
    for i in range(count):
        W, h = host_layers[i % 3]
        dst = out_hosts[i % nstreams]

        def to_host(run_, kept, dst=dst):
            dst["Q_packed"].copy_(run_.Q_packed, non_blocking=True)
            dst["L"].copy_(run_.L, non_blocking=True)
            dst["R"].copy_(run_.R, non_blocking=True)
        t0 = time.perf_counter()
        hd = caldera_async(qp, W, h, device=dev, use_tqdm=False, W_copy="none", seed=1000, return_dense=False,
                           return_packed=False, consume=to_host, slots=a.slots, batch=a.batch)
        g = hd._group
        if hd._index == 0 and keep and not g.launched:
            rec = {"slot": g.slot.index, "first": torch.cuda.Event(enable_timing=True), "t_first": t0}
            rec["first"].record(g.slot.stream)
            g._trace = rec
            trace.append(rec)
        t1 = time.perf_counter()
        pending.append(hd)
        if len(pending) > nstreams:
            pending.pop(0).result()
        t2 = time.perf_counter()
        calls.append((t1 - t0, t2 - t1, t0))
    engine.flush()
    for hd in pending:
        hd.result()
    torch.cuda.synchronize()
    return calls


run(2 * nstreams, False)
origin.record()
t_begin = time.perf_counter()
calls = run(a.steps * nstreams, True)
total = time.perf_counter() - t_begin
print(f"{a.slots}x{a.batch}: {a.steps * nstreams / total:.1f} matrices/s over {total:.3f} s")
sub = sorted(c[0] for c in calls)
har = sorted(c[1] for c in calls)
print(f"caldera_async() host time: median {sub[len(sub) // 2] * 1e3:.2f} ms, p99 {sub[int(len(sub) * 0.99)] * 1e3:.2f} ms, max {sub[-1] * 1e3:.1f} ms, "
      f"sum {sum(sub):.3f} s;  result() wait: max {har[-1] * 1e3:.1f} ms, sum {sum(har):.3f} s")
print("batch slot  first_copy_done@ms  staged@ms  done@ms   copies_ms  compute_ms  host_submit_span_ms  host_first@ms  host_launch_call_ms")
for k, rec in enumerate(trace):
    if "done" not in rec:
        continue
    f = origin.elapsed_time(rec["first"])
    s = origin.elapsed_time(rec["staged"])
    d = origin.elapsed_time(rec["done"])
    print(f"{k:4d} {rec['slot']:4d}  {f:10.1f} {s:10.1f} {d:10.1f}   {s - f:8.1f} {d - s:9.1f}   {(rec['t_launch'] - rec['t_first']) * 1e3:8.1f}   {(rec['t_first'] - t_begin) * 1e3:9.1f} {(rec['t_launched'] - rec['t_launch']) * 1e3:9.1f}")
release_engines()

print("\nhost-side events longer than 15 ms (ms since the start of the timed run):")
events = []
for sub, har, t0 in calls:
    if sub > 0.015:
        events.append((t0 - t_begin, f"submit thread: caldera_async() took {sub * 1e3:.1f} ms"))
    if har > 0.015:
        events.append((t0 + sub - t_begin, f"submit thread: result() took {har * 1e3:.1f} ms"))
for slot, a0, r0, r1, a1 in issue_log:
    if a0 >= t_begin and a1 - a0 > 0.015:
        events.append((a0 - t_begin, f"launcher of slot {slot}: issue took {(a1 - a0) * 1e3:.1f} ms, of which graph replay call "
                                     f"{(r1 - r0) * 1e3:.1f} ms, consume hooks + record {(a1 - r1) * 1e3:.1f} ms"))
for t, msg in sorted(events):
    print(f"  {t * 1e3:8.1f}  {msg}")
