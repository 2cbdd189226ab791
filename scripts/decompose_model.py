"""Full-model job (BASELINE config 4): decompose every linear layer of a random-init
Llama-2-7B-shaped model (224 layers: 32 x {4 x 4096x4096, 2 x 11008x4096, 1 x 4096x11008}),
layer-sharded over the GPUs of one box, and gather the packed results on rank 0.

    python scripts/decompose_model.py [--layers-per-block 7 --blocks 32 --rank 128 --lbits 16]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/decompose_model.py

Prints one JSON line per pass with the wall-clock seconds of the decomposition and of the gather (max over
ranks, device synchronised, weights generated on the owning GPU outside the timed region); pass 0 includes
the one-time graph captures."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ee274_convexcaldera_llm_quantization_b200 import model_job as mj  # noqa: E402
from ee274_convexcaldera_llm_quantization_b200 import scheduler as sch  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams  # noqa: E402
from src.caldera.utils.quantization import QuantizerFactory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=32)
ap.add_argument("--rank", type=int, default=128)
ap.add_argument("--lbits", type=int, default=16)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--streams", type=int, default=48)
ap.add_argument("--repeats", type=int, default=2, help="timed passes after the warm-up pass")
ap.add_argument("--hidden", type=int, default=4096)
ap.add_argument("--ffn", type=int, default=11008)
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

names, shapes = mj.llama_shapes(a.blocks, a.hidden, a.ffn)
fac = QuantizerFactory(method="uniform", block_size=64)
params = CalderaParams(Q_bits=2, L_bits=a.lbits, R_bits=a.lbits, rank=a.rank, iters=a.iters, lplr_iters=5,
                       update_order=["Q", "LR"], quant_factory_Q=fac, quant_factory_LR=fac)
shards = sch.shard_layout(params, shapes, world)[0]
store = mj.synth_layers(shapes, shards[rank], dev)      # generated on the owning GPU before the clock starts


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


sch.warm_up_gather(dev, dst=0)
best = None
for attempt in range(1 + a.repeats):                     # pass 0 captures the graphs of every shape in every slot
    out = mj.run_model_job(params, names, shapes, store, rank, world, dev, streams=a.streams, barrier=barrier)
    tt = torch.tensor([out["decompose_s"], out["gather_s"], out["wall_s"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        parts = sch.split_gathered(out["arena"], out["shards"], out["sizes"])
        first = sch.unpack_decomposition(parts[0])
        line = {"job": "llama2-7b-shape decomposition", "pass": attempt, "layers": len(parts),
                "params": sum(m * n for m, n in shapes), "n_gpus": world, "rank": a.rank, "L_R_bits": a.lbits,
                "iters": a.iters, "streams": a.streams, "decompose_wall_s": float(tt[0]), "gather_wall_s": float(tt[1]),
                "wall_s": float(tt[2]), "gathered_bytes": int(out["gathered_bytes"]),
                "layers_per_s": len(parts) / float(tt[2]), "first_layer": first["name"],
                "first_layer_best_error": min(first["errors"]["LR"])}
        print(json.dumps(line), flush=True)
    del out
if world > 1:
    dist.destroy_process_group()
