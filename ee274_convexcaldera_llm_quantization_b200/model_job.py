"""The model-level job of BASELINE config 4: every linear layer of a Llama-2-7B-shaped model decomposed,
layer-sharded over the ranks of one box, packed results gathered on one rank.

Replaces the sequential per-layer loop of the reference's driver (main.py:147-199) for the shapes the north star
names: 32 blocks x {q,k,v,o: 4096 x 4096; gate,up: 11008 x 4096; down: 4096 x 11008} = 224 layers, 6.476 G
parameters.  Used by bench.py (the `full_7b_wall_s` keys) and scripts/decompose_model.py."""
from __future__ import annotations

import time
from typing import Dict, List, Sequence, Tuple

import torch

from . import scheduler as sch


def llama_shapes(blocks: int = 32, hidden: int = 4096, ffn: int = 11008) -> Tuple[List[str], List[Tuple[int, int]]]:
    names, shapes = [], []
    for b in range(blocks):
        for nm, shp in (("q_proj", (hidden, hidden)), ("k_proj", (hidden, hidden)), ("v_proj", (hidden, hidden)),
                        ("o_proj", (hidden, hidden)), ("gate_proj", (ffn, hidden)), ("up_proj", (ffn, hidden)),
                        ("down_proj", (hidden, ffn))):
            names.append(f"layers.{b}.{nm}")
            shapes.append(shp)
    return names, shapes


def synth_layers(shapes: Sequence[Tuple[int, int]], indices: Sequence[int], dev: torch.device) -> Dict[int, tuple]:
    """SURVEY 8d synthetic inputs of the layers a rank owns, generated on its GPU (seed 1000 + layer index):
    W = 0.02 N(0, 1), h = 0.5 + U(0, 1)."""
    store = {}
    for i in indices:
        g = torch.Generator(device=dev).manual_seed(1000 + i)
        m, n = shapes[i]
        store[i] = (0.02 * torch.randn(m, n, generator=g, device=dev), 0.5 + torch.rand(n, generator=g, device=dev))
    return store


def run_model_job(params, names, shapes, store, rank: int, world: int, dev: torch.device, streams: int = 48,
                  dst: int = 0, factor_dtype: str = "float16", barrier=None, slots: int = 3) -> dict:
    """One timed pass: decompose this rank's shard (inputs already on its GPU) and gather the packed blobs on `dst`.
    Wall clock from a barrier to the end of the gather, device synchronised on both sides; the caller takes the
    max over ranks.  Returns the timings, the gathered arena (on `dst`) and this rank's ShardResult."""
    shards, sizes, rank_bytes, rank_offs = sch.shard_layout(params, shapes, world, factor_dtype)
    layers = [(names[i], (lambda i=i: store[i])) for i in range(len(names))]
    total = sum(rank_bytes)
    with torch.cuda.device(dev):
        # the destination rank decomposes straight into its slice of the gather target
        big = torch.zeros(total, dtype=torch.uint8, device=dev) if rank == dst else None
        arena = big[rank_offs[rank]:rank_offs[rank] + rank_bytes[rank]] if big is not None else \
            torch.zeros(max(rank_bytes[rank], 1), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        if barrier is not None:
            barrier()
        from .engine import get_engine
        captured = get_engine(dev, slots, max(1, streams // max(1, slots))).captures
        t0 = time.perf_counter()
        shard = sch.decompose_layers(layers, shapes, params, rank, world, device=dev, streams=streams, slots=slots,
                                     factor_dtype=factor_dtype, arena=arena)
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        if world > 1:
            big = sch.gather_arena(arena, rank_bytes, dst=dst, out=big)
            torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
    return {"decompose_s": t1 - t0, "gather_s": t2 - t1, "wall_s": t2 - t0, "gathered_bytes": total, "arena": big,
            "graphs_captured": get_engine(dev).captures - captured,
            "shard": shard, "shards": shards, "sizes": sizes}
