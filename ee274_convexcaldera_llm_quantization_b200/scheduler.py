"""Layer-sharded multi-GPU scheduler.

The reference decomposes a model's linear layers strictly sequentially on one device
(main.py:147-199).  Layers are independent, so here every rank (one process per GPU)
decomposes a disjoint, cost-balanced subset with no communication on the critical path;
the only collective is one gather of the packed results at the end (NCCL over NVLink when
the blobs live on the GPU, gloo in the CPU tests).

  lpt_assign          longest-processing-time greedy partition of layers over ranks
  pack_decomposition  CalderaDecomposition -> one flat uint8 blob (packed codes, scales, factors)
  unpack_decomposition
  gather_blobs        variable-length gather of per-layer blobs to one rank (or all ranks)
  decompose_layers    run caldera() over this rank's shard
"""
from __future__ import annotations

import json
import os
import time
import struct
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

_MAGIC = b"CALDB200"


# ----------------------------------------------------------------------------- partitioning
def layer_cost(m: int, n: int, rank: int, iters: int, lplr_iters: int = 0, quantised_factors: bool = False,
               sketch_width: Optional[int] = None, power_iters: int = 12, power_iters_warm: int = 3) -> float:
    """Flop model of one layer: sketch passes dominate.  The first rank-r step starts
    cold (`power_iters`), the later ones from the previous basis (`power_iters_warm`)."""
    q = sketch_width if sketch_width else max(2 * rank, rank + 32)
    it = max(iters, 1)
    sketch = 2.0 * m * n * q * ((2 + 2 * power_iters) + (it - 1) * (2 + 2 * power_iters_warm))
    rest = 4.0 * m * n * rank
    if quantised_factors:
        rest += 6.0 * m * n * rank * lplr_iters
    return sketch + rest * it


def lpt_assign(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy LPT: heaviest layer first onto the least loaded rank.  Deterministic (ties by
    index), so every rank computes the same map without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    for s in shards:
        s.sort()
    return shards


# ----------------------------------------------------------------------------- wire format
_FIELDS = ("Q_packed", "Q_scale", "L", "R", "L_packed", "R_packed", "L_scale", "R_scale")
HEADER_BYTES = 4096          # fixed-size header region at the head of every job blob (magic | length | JSON | padding)
_FACTOR_DTYPES = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


def blob_layout(params, shape: Tuple[int, int], factor_dtype: str = "float16"):
    """Byte layout of one layer's blob: a pure function of the layer shape and the parameters, so every rank knows
    every other rank's blob sizes without exchanging them.  Returns (total_bytes, [(field, dtype, shape, offset,
    nbytes)]), offsets relative to the blob's start (the payload begins at HEADER_BYTES)."""
    from . import _lib
    lib = _lib.load()
    m, n = int(shape[0]), int(shape[1])
    r = int(params.rank)
    secs, off = [], HEADER_BYTES

    def add(field, dtype, shp, nbytes):
        nonlocal off
        secs.append((field, dtype, tuple(shp), off, int(nbytes)))
        off += _align(int(nbytes), 16)
    if params.compute_quantized_component:
        nb = int(lib.cb_packed_bytes(m * n, int(params.Q_bits)))
        add("Q_packed", "uint8", (nb,), nb)
        add("Q_scale", "float32", (1,), 4)
    if params.compute_low_rank_factors:
        if params.L_bits < 16 or params.R_bits < 16:
            # either factor below 16 bits sends BOTH through the quantiser (alg.py:144): codes of L_bits / R_bits each
            nl, nr = int(lib.cb_packed_bytes(m * r, int(params.L_bits))), int(lib.cb_packed_bytes(r * n, int(params.R_bits)))
            add("L_packed", "uint8", (nl,), nl)
            add("R_packed", "uint8", (nr,), nr)
            add("L_scale", "float32", (1,), 4)
            add("R_scale", "float32", (1,), 4)
        else:
            es = torch.empty(0, dtype=_FACTOR_DTYPES[factor_dtype]).element_size()
            add("L", factor_dtype, (m, r), m * r * es)
            add("R", factor_dtype, (r, n), r * n * es)
    return _align(off), secs


def _header_bytes(meta: dict) -> bytes:
    header = json.dumps(meta).encode()
    if len(_MAGIC) + 8 + len(header) > HEADER_BYTES:
        # an error trajectory this long does not fit the fixed header: keep its tail
        meta = dict(meta)
        meta["errors"] = {k: v[-8:] for k, v in meta["errors"].items()}
        meta["errors_truncated"] = True
        header = json.dumps(meta).encode()
    pad = HEADER_BYTES - len(_MAGIC) - 8 - len(header)
    return _MAGIC + struct.pack("<Q", len(header) + pad) + header + b" " * pad


def pack_decomposition(name: str, dec, q_bits: int, l_bits: int, r_bits: int, shape: Tuple[int, int],
                       out: Optional[torch.Tensor] = None, factor_dtype: str = "float16") -> torch.Tensor:
    """Serialises what a consumer of the decomposition needs: packed Q codes + scale, and the factors (packed codes
    + scales when quantised; `factor_dtype` -- float16 by default, bfloat16 or float32 -- otherwise).
    Layout: magic | u64 header length | JSON header | 16-byte aligned payload sections.
    `out` (uint8, device): write the blob into its head and return that view when it fits."""
    tensors: Dict[str, torch.Tensor] = {}
    for f in _FIELDS:
        t = getattr(dec, f, None)
        if torch.is_tensor(t):
            tensors[f] = t
    if "L_packed" in tensors:       # quantised factors travel as codes, not as dense fp32
        tensors.pop("L", None)
        tensors.pop("R", None)
    else:
        for f in ("L", "R"):
            if f in tensors:
                tensors[f] = tensors[f].to(_FACTOR_DTYPES[factor_dtype])
    meta = {"name": name, "shape": list(shape), "q_bits": q_bits, "l_bits": l_bits, "r_bits": r_bits,
            "global_scale": float(dec.global_scale), "best_step": int(getattr(dec, "best_step", -1)),
            "errors": dec.errors, "sections": []}
    device = next(iter(tensors.values())).device if tensors else torch.device("cpu")
    parts, offset = [], 0
    for f, t in tensors.items():
        raw = t.contiguous().reshape(-1).view(torch.uint8)
        pad = (-raw.numel()) % 16
        meta["sections"].append({"field": f, "dtype": str(t.dtype).replace("torch.", ""), "shape": list(t.shape),
                                 "offset": offset, "nbytes": raw.numel()})
        parts.append(raw)
        if pad:
            parts.append(torch.zeros(pad, dtype=torch.uint8, device=raw.device))
        offset += raw.numel() + pad
    header = json.dumps(meta).encode()
    hpad = (-(len(_MAGIC) + 8 + len(header))) % 16
    head = _MAGIC + struct.pack("<Q", len(header) + hpad) + header + b" " * hpad
    head_t = torch.frombuffer(bytearray(head), dtype=torch.uint8).to(device)
    total = head_t.numel() + offset
    if out is not None and out.numel() >= total and out.device == head_t.device:
        return torch.cat([head_t] + parts, out=out[:total])
    return torch.cat([head_t] + parts) if parts else head_t


def unpack_decomposition(blob: torch.Tensor) -> dict:
    raw = blob.detach().cpu().contiguous()
    b = raw[:16].numpy().tobytes()
    if b[:8] != _MAGIC:
        raise ValueError("not a caldera-b200 blob")
    (hlen,) = struct.unpack("<Q", b[8:16])
    meta = json.loads(raw[16:16 + hlen].numpy().tobytes().decode())
    base = 16 + hlen
    out = dict(meta)
    for sec in meta["sections"]:
        dt = getattr(torch, sec["dtype"])
        seg = raw[base + sec["offset"]: base + sec["offset"] + sec["nbytes"]].clone()
        out[sec["field"]] = seg.view(dt).reshape(sec["shape"])
    return out


# ----------------------------------------------------------------------------- gather
def shard_layout(params, shapes: Sequence[Tuple[int, int]], world_size: int, factor_dtype: str = "float16"):
    """What every rank can compute on its own: the layer -> rank map and the byte offsets of every layer's blob
    inside its rank's arena and inside the gathered arena (ranks in order, layers in index order within a rank).
    Returns (shards, sizes, rank_bytes, rank_offsets)."""
    quantised = params.compute_low_rank_factors and (params.L_bits < 16 or params.R_bits < 16)
    costs = [layer_cost(m, n, params.rank, params.iters, params.lplr_iters, quantised) for (m, n) in shapes]
    shards = lpt_assign(costs, world_size)
    sizes = [blob_layout(params, shp, factor_dtype)[0] for shp in shapes]
    rank_bytes = [sum(sizes[i] for i in sh) for sh in shards]
    rank_offsets = [sum(rank_bytes[:r]) for r in range(world_size)]
    return shards, sizes, rank_bytes, rank_offsets


def warm_up_gather(device: torch.device, dst: Optional[int] = 0, group=None) -> None:
    """Sets up the point-to-point channels the gather uses (NCCL creates them lazily on first use, which costs
    far more than the gather itself) by exchanging one tiny message along every edge."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    tiny = torch.zeros(world * 256, dtype=torch.uint8, device=device)
    gather_arena(tiny[rank * 256:(rank + 1) * 256], [256] * world, dst=dst, group=group, out=tiny)
    if device.type == "cuda":
        torch.cuda.synchronize(device)


def gather_arena(mine: torch.Tensor, rank_bytes: Sequence[int], dst: Optional[int] = 0, group=None,
                 out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """The one collective of the job: every rank's arena (uint8, rank_bytes[rank] bytes, sizes known to all) lands
    at its offset in ONE preallocated arena on `dst` (dst=None: on every rank).  Receives are posted straight into
    their final position -- no padding to the largest shard, no per-rank staging buffers, no size exchange.  Runs
    as a single group of point-to-point operations (NCCL over NVLink for device arenas, gloo for CPU tensors)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    offs = [sum(rank_bytes[:r]) for r in range(world)]
    total = sum(rank_bytes)
    receive = dst is None or rank == dst
    if receive:
        if out is None:
            out = torch.empty(total, dtype=torch.uint8, device=mine.device)
        own = out[offs[rank]:offs[rank] + rank_bytes[rank]]
        if own.data_ptr() != mine.data_ptr() and rank_bytes[rank] > 0:
            own.copy_(mine[:rank_bytes[rank]])
    if dst is None:
        # every shard is broadcast from its owner into place (world collectives at full fabric bandwidth)
        works = [dist.broadcast(out[offs[r]:offs[r] + rank_bytes[r]], src=dist.get_global_rank(group, r) if group else r,
                                group=group, async_op=True) for r in range(world) if rank_bytes[r] > 0]
        for w in works:
            w.wait()
        return out
    ops = []
    if rank == dst:
        for r in range(world):
            if r != rank and rank_bytes[r] > 0:
                ops.append(dist.P2POp(dist.irecv, out[offs[r]:offs[r] + rank_bytes[r]], r, group))
    elif rank_bytes[rank] > 0:
        ops.append(dist.P2POp(dist.isend, mine[:rank_bytes[rank]], dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out if receive else None


def split_gathered(arena: torch.Tensor, shards: Sequence[Sequence[int]], sizes: Sequence[int]) -> Dict[int, torch.Tensor]:
    """layer index -> blob view inside a gathered arena (see shard_layout)."""
    out, off = {}, 0
    for sh in shards:
        for i in sh:
            out[i] = arena[off:off + sizes[i]]
            off += sizes[i]
    return out


def gather_blobs(blobs: List[torch.Tensor], dst: Optional[int] = 0, group=None) -> Optional[List[List[torch.Tensor]]]:
    """Gathers every rank's list of variable-length blobs when the sizes are NOT known in advance (free-form blobs
    from pack_decomposition).  dst=None -> all ranks receive.  One small all_gather of the sizes, then
    `gather_arena`.  Returns, on receiving ranks, result[rank] = list of that rank's blobs; None elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if blobs:
        device = blobs[0].device
    elif dist.get_backend(group) == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    else:
        device = torch.device("cpu")
    sizes_obj: List[Optional[List[int]]] = [None] * world
    dist.all_gather_object(sizes_obj, [int(b.numel()) for b in blobs], group=group)
    rank_bytes = [sum(sz) for sz in sizes_obj]
    mine = torch.cat(blobs) if blobs else torch.empty(0, dtype=torch.uint8, device=device)
    arena = gather_arena(mine, rank_bytes, dst=dst, group=group)
    if arena is None:
        return None
    out: List[List[torch.Tensor]] = []
    off = 0
    for r in range(world):
        items = []
        for sz in sizes_obj[r]:
            items.append(arena[off:off + sz])
            off += sz
        out.append(items)
    return out


# ----------------------------------------------------------------------------- driver
class ShardResult:
    """This rank's part of a job: `indices` (layer indices, ascending), `arena` (uint8 device tensor holding their
    blobs back to back in that order), `blobs` (views), `records` (per-layer dict: errors, global_scale, ...)."""

    def __init__(self, indices, arena, blobs, records):
        self.indices, self.arena, self.blobs, self.records = indices, arena, blobs, records

    def __iter__(self):          # (indices, results) like the round-1 return value
        return iter((self.indices, self.blobs))


def decompose_layers(layers: Sequence[Tuple[str, Callable[[], Tuple[torch.Tensor, Optional[torch.Tensor]]]]],
                     shapes: Sequence[Tuple[int, int]], params, rank: int, world_size: int,
                     device: Optional[torch.device] = None, pack: bool = True, streams: int = 48,
                     factor_dtype: str = "float16", arena: Optional[torch.Tensor] = None, slots: int = 3,
                     **caldera_kwargs):
    """Decomposes this rank's shard of `layers` (the loop of main.py:147-199, layer-sharded and asynchronous).

    layers[i] = (name, loader) where loader() returns (W, H) -- generated or loaded directly on the owning GPU (or
    in pinned host memory), so no weight ever crosses ranks.  ONE host thread submits every layer to the device's
    LayerEngine (engine.py), which keeps up to `streams` of them in flight -- `slots` graph replays of `streams /
    slots` same-shape layers advancing in lock step (largest layers first) -- and never blocks on a layer's result;
    the packed outputs are copied device-side straight into this rank's wire-format arena and only the ~100-byte
    result records come back to the host, to be written into the blob headers at the end.

    pack=True returns a ShardResult (unpacks as (indices, blobs)); pack=False returns (indices, decompositions).
    The per-layer seed is derived from the layer index only and no kernel on the path uses floating-point atomics
    for its outputs, so the sharded run equals the single-GPU run bit for bit.  The job runs in the library's
    "throughput" execution mode whatever `streams` and the world size are; the previous mode is restored on return.
    `arena`: optional preallocated uint8 device tensor for this rank's blobs (e.g. a slice of the gather target on
    the destination rank, which makes its own contribution zero-copy)."""
    from . import _lib
    previous_mode = _lib.execution_mode()
    _lib.set_execution_mode("throughput")
    try:
        return _decompose_layers(layers, shapes, params, rank, world_size, device, pack, streams, factor_dtype, arena,
                                 slots, caldera_kwargs)
    finally:
        _lib.set_execution_mode(previous_mode)


def batch_plan(count: int, batch: int) -> List[int]:
    """Sizes of the groups `count` same-shape layers run in when at most `batch` advance together: as few groups as
    possible, sizes differing by at most one."""
    if count <= 0:
        return []
    k = -(-count // max(1, batch))
    base, rem = divmod(count, k)
    return [base + 1] * rem + [base] * (k - rem)


def _decompose_layers(layers, shapes, params, rank, world_size, device, pack, streams, factor_dtype, arena, slots,
                      caldera_kwargs):
    from .alg import caldera_async
    from .engine import get_engine
    shards, sizes, rank_bytes, _ = shard_layout(params, shapes, world_size, factor_dtype)
    mine = shards[rank]
    quantised = params.compute_low_rank_factors and (params.L_bits < 16 or params.R_bits < 16)
    costs = {i: layer_cost(shapes[i][0], shapes[i][1], params.rank, params.iters, params.lplr_iters, quantised) for i in mine}
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    order = sorted(mine, key=lambda i: (-costs[i], shapes[i], i))          # big layers first, equal shapes together
    slots = max(1, int(slots))
    batch = max(1, int(streams) // slots)
    engine = get_engine(dev, slots, batch)
    # a captured batch always advances all of its layers, so a shape's layers are split into equal groups up front
    # (17 layers at batch 16 -> 9 + 8) instead of ending in a mostly empty batch
    per_shape, group_of = {}, {}
    for i in order:
        per_shape.setdefault(tuple(shapes[i]), []).append(i)
    for members in per_shape.values():
        sizes_ = batch_plan(len(members), batch)
        pos = 0
        for sz in sizes_:
            for i in members[pos:pos + sz]:
                group_of[i] = sz
            pos += sz
    t_start = time.perf_counter()
    with torch.cuda.device(dev):
        offsets, off = {}, 0
        for i in mine:
            offsets[i] = off
            off += sizes[i]
        if pack:
            if arena is None:
                arena = torch.zeros(max(off, 1), dtype=torch.uint8, device=dev)   # padding bytes are part of the blob
            assert arena.numel() >= off and arena.dtype == torch.uint8
        layouts = {shp: blob_layout(params, shp, factor_dtype)[1] for shp in {shapes[i] for i in mine}}
        fdt = _FACTOR_DTYPES[factor_dtype]
        handles = {}
        for i in order:
            name, loader = layers[i]
            W, H = loader()
            kw = dict(caldera_kwargs)
            kw.setdefault("seed", 1000 + i)
            kw.setdefault("W_copy", "none")
            if pack:
                blob = arena[offsets[i]:offsets[i] + sizes[i]]

                def consume(run, kept, blob=blob, secs=layouts[tuple(W.shape)]):
                    # device-side copies of the layer's outputs into its blob (slot stream current)
                    for field, dtype, shp, o, nb in secs:
                        src = getattr(run, field)
                        dst = blob[o:o + nb]
                        if field in ("L", "R"):
                            dst.view(fdt).reshape(shp).copy_(src)       # fp32 -> wire dtype
                            if fdt == torch.float16:
                                kept["finite"] = kept.get("finite", True) & torch.isfinite(dst.view(fdt)).all()
                        else:
                            dst.copy_(src.reshape(-1).view(torch.uint8))
                kw.update(return_dense=False, return_packed=False, consume=consume)
            handles[i] = caldera_async(params, W, H, device=dev, use_tqdm=False, slots=slots, batch=batch,
                                       batch_hint=group_of[i], **kw)
        engine.flush()
        t_submit = time.perf_counter()
        results, records = {}, {}
        for i in mine:
            dec = handles[i].result()
            if pack:
                fin = handles[i].kept.get("finite")
                if fin is not None and not bool(fin):
                    raise OverflowError(f"layer {layers[i][0]}: a 16-bit factor does not fit float16 on the wire "
                                        "(pass factor_dtype='bfloat16' or 'float32')")
                meta = {"name": layers[i][0], "shape": list(shapes[i]), "q_bits": params.Q_bits, "l_bits": params.L_bits,
                        "r_bits": params.R_bits, "global_scale": float(dec.global_scale), "best_step": int(dec.best_step),
                        "errors": dec.errors,
                        "sections": [{"field": f, "dtype": d, "shape": list(shp), "offset": o - HEADER_BYTES, "nbytes": nb}
                                     for f, d, shp, o, nb in layouts[tuple(shapes[i])]]}
                records[i] = meta
                results[i] = arena[offsets[i]:offsets[i] + sizes[i]]
            else:
                results[i] = dec
        if pack and mine:
            # every header in one pinned staging buffer and one strided device copy
            heads = torch.empty((len(mine), HEADER_BYTES), dtype=torch.uint8).pin_memory()
            for j, i in enumerate(mine):
                heads[j] = torch.frombuffer(bytearray(_header_bytes(records[i])), dtype=torch.uint8)
            heads_d = heads.to(dev, non_blocking=True)
            for j, i in enumerate(mine):
                arena[offsets[i]:offsets[i] + HEADER_BYTES].copy_(heads_d[j], non_blocking=True)
            torch.cuda.current_stream().synchronize()
    if os.environ.get("CB_SCHEDULER_TIMES"):
        import sys
        sys.stderr.write(f"[decompose_layers] rank {rank}: {len(mine)} layers, submit {t_submit - t_start:.3f} s, "
                         f"total {time.perf_counter() - t_start:.3f} s, graphs captured so far {engine.captures}, "
                         f"groups { {k: batch_plan(len(v), batch) for k, v in per_shape.items()} }\n")
    if pack:
        return ShardResult(mine, arena, [results[i] for i in mine], records)
    return mine, [results[i] for i in mine]
