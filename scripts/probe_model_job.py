"""The 224-layer Llama-2-7B-shape job on one GPU, several passes, with the scheduler's own timing lines
(CB_SCHEDULER_TIMES=1): how long submission takes, how many graphs a pass captures, pass-to-pass variance.

  python scripts/probe_model_job.py [--lbits 16] [--passes 4] [--slots 3] [--streams 48]
"""
import argparse
import os
import sys
import time

os.environ.setdefault("CB_SCHEDULER_TIMES", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ee274_convexcaldera_llm_quantization_b200 import model_job as mj, scheduler as sch  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams  # noqa: E402
from src.caldera.utils.quantization import QuantizerFactory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lbits", type=int, default=16)
ap.add_argument("--passes", type=int, default=4)
ap.add_argument("--slots", type=int, default=3)
ap.add_argument("--streams", type=int, default=48)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:                      # under torchrun: the sharded job with its gather, per-rank times printed by every rank
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
fac = QuantizerFactory(method="uniform", block_size=64)
prm = CalderaParams(Q_bits=2, L_bits=a.lbits, R_bits=a.lbits, rank=128, iters=5, lplr_iters=5, activation_aware_LR=True,
                    update_order=["Q", "LR"], quant_factory_Q=fac, quant_factory_LR=fac, rand_svd=False, sigma_reg=0)
names, shapes = mj.llama_shapes(32)
shards = sch.shard_layout(prm, shapes, world)[0]
store = mj.synth_layers(shapes, shards[rank], dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


if world > 1:
    sch.warm_up_gather(dev, dst=0)
res = None
for k in range(a.passes):
    res = None
    ms0 = torch.cuda.memory_stats(dev)
    t0 = time.perf_counter()
    res = mj.run_model_job(prm, names, shapes, store, rank, world, dev, streams=a.streams, slots=a.slots, barrier=barrier)
    ms1 = torch.cuda.memory_stats(dev)
    print(f"rank {rank} pass {k}: wall {res['wall_s']:.3f} s decompose {res['decompose_s']:.3f} gather {res['gather_s']:.4f} "
          f"(call {time.perf_counter() - t0:.3f} s), graphs captured {res['graphs_captured']}, cudaMalloc/cudaFree "
          f"{ms1.get('num_device_alloc', 0) - ms0.get('num_device_alloc', 0)}/{ms1.get('num_device_free', 0) - ms0.get('num_device_free', 0)}, "
          f"reserved {ms1.get('reserved_bytes.all.current', 0) / 2**30:.1f} GiB", flush=True)
if world > 1:
    dist.destroy_process_group()
