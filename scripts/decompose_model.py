"""Full-model job (BASELINE config 4): decompose every linear layer of a random-init
Llama-2-7B-shaped model (224 layers: 32 x {4 x 4096x4096, 2 x 11008x4096, 1 x 4096x11008}),
layer-sharded over the GPUs of one box, and gather the packed results on rank 0.

    python scripts/decompose_model.py [--layers-per-block 7 --blocks 32 --rank 128 --lbits 16]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/decompose_model.py

Prints one JSON line with the wall-clock seconds of the decomposition (max over ranks, device
synchronised, weights generated on the owning GPU outside the timed region) and of the gather."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ee274_convexcaldera_llm_quantization_b200 import scheduler as sch  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams  # noqa: E402
from src.caldera.utils.quantization import QuantizerFactory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=32)
ap.add_argument("--rank", type=int, default=128)
ap.add_argument("--lbits", type=int, default=16)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--streams", type=int, default=16)
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

names, shapes = [], []
for b in range(a.blocks):
    for nm, shp in (("q_proj", (a.hidden, a.hidden)), ("k_proj", (a.hidden, a.hidden)), ("v_proj", (a.hidden, a.hidden)),
                    ("o_proj", (a.hidden, a.hidden)), ("gate_proj", (a.ffn, a.hidden)), ("up_proj", (a.ffn, a.hidden)),
                    ("down_proj", (a.hidden, a.ffn))):
        names.append(f"layers.{b}.{nm}")
        shapes.append(shp)
fac = QuantizerFactory(method="uniform", block_size=64)
params = CalderaParams(Q_bits=2, L_bits=a.lbits, R_bits=a.lbits, rank=a.rank, iters=a.iters, lplr_iters=5,
                       update_order=["Q", "LR"], quant_factory_Q=fac, quant_factory_LR=fac)
quantised = a.lbits < 16
costs = [sch.layer_cost(m, n, a.rank, a.iters, 5, quantised) for m, n in shapes]
mine = sch.lpt_assign(costs, world)[rank]

# synthetic weights of the owning rank, generated on its GPU before the clock starts (SURVEY 8d seeds)
store = {}
for i in mine:
    g = torch.Generator(device=dev).manual_seed(1000 + i)
    m, n = shapes[i]
    store[i] = (0.02 * torch.randn(m, n, generator=g, device=dev), 0.5 + torch.rand(n, generator=g, device=dev))
layers = [(names[i], (lambda i=i: store[i])) for i in range(len(names))]

# warm-up: one small-rank pass per distinct shape builds workspaces, graphs and attributes
# (`streams` concurrent copies, so that every worker finds a captured graph in the pool)
for shp in sorted(set(shapes[i] for i in mine)):
    i = next(k for k in mine if shapes[k] == shp)
    sch.decompose_layers([layers[i]] * a.streams, [shapes[i]] * a.streams, params, 0, 1, device=dev, streams=a.streams)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
idx, blobs = sch.decompose_layers(layers, shapes, params, rank, world, device=dev, streams=a.streams)
torch.cuda.synchronize()
t_dec = time.perf_counter() - t0
t1 = time.perf_counter()
gathered = sch.gather_blobs(blobs, dst=0) if world > 1 else [blobs]
torch.cuda.synchronize()
t_gather = time.perf_counter() - t1
tt = torch.tensor([t_dec, t_gather], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    nbytes = sum(b.numel() for r in gathered for b in r)
    nlayers = sum(len(r) for r in gathered)
    first = sch.unpack_decomposition(gathered[0][0])
    print(json.dumps({"job": "llama2-7b-shape decomposition", "layers": nlayers, "params": sum(m * n for m, n in shapes),
                      "n_gpus": world, "rank": a.rank, "L_R_bits": a.lbits, "iters": a.iters, "streams": a.streams,
                      "decompose_wall_s": float(tt[0]), "gather_wall_s": float(tt[1]), "gathered_bytes": int(nbytes),
                      "layers_per_s": nlayers / float(tt[0]), "first_layer": first["name"],
                      "first_layer_best_error": min(first["errors"]["LR"])}))
if os.environ.get("CB_CALDERA_TIMES"):
    from ee274_convexcaldera_llm_quantization_b200.alg import phase_times
    print(f"rank {rank} host phases of caldera() summed over worker threads: {phase_times()}", file=sys.stderr)
if world > 1:
    dist.destroy_process_group()
