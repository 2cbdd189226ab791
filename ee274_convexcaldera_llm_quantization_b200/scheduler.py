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
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

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


def pack_decomposition(name: str, dec, q_bits: int, l_bits: int, r_bits: int, shape: Tuple[int, int],
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Serialises what a consumer of the decomposition needs: packed Q codes + scale, and the
    factors (packed codes + scales when quantised, fp16 otherwise; fp32 if `l_bits` >= 32).
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
                tensors[f] = tensors[f].to(torch.float16)
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
    b = raw.numpy().tobytes()
    if b[:8] != _MAGIC:
        raise ValueError("not a caldera-b200 blob")
    (hlen,) = struct.unpack("<Q", b[8:16])
    meta = json.loads(b[16:16 + hlen].decode())
    base = 16 + hlen
    out = dict(meta)
    for sec in meta["sections"]:
        dt = getattr(torch, sec["dtype"])
        seg = raw[base + sec["offset"]: base + sec["offset"] + sec["nbytes"]].clone()
        out[sec["field"]] = seg.view(dt).reshape(sec["shape"])
    return out


def gather_blobs(blobs: List[torch.Tensor], dst: Optional[int] = 0, group=None) -> Optional[List[List[torch.Tensor]]]:
    """Gathers every rank's list of blobs.  dst=None -> all ranks receive (all_gather).
    Two collectives: blob sizes, then one padded payload per rank.  Returns, on receiving
    ranks, result[rank] = list of that rank's blobs; None elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if blobs:
        device = blobs[0].device
    elif dist.get_backend(group) == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    else:
        device = torch.device("cpu")
    sizes = torch.tensor([b.numel() for b in blobs], dtype=torch.int64, device=device)
    count = torch.tensor([len(blobs)], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts_host = [int(c) for c in torch.cat(counts).tolist()]          # one device read instead of one per rank
    max_count = max(counts_host)
    size_pad = torch.zeros(max(max_count, 1), dtype=torch.int64, device=device)
    size_pad[:len(blobs)] = sizes
    all_sizes = [torch.zeros_like(size_pad) for _ in range(world)]
    dist.all_gather(all_sizes, size_pad, group=group)
    sizes_host = torch.stack(all_sizes).tolist()                        # [rank][k], one device read
    totals = [int(sum(row)) for row in sizes_host]
    max_total = max(max(totals), 1)
    payload = torch.zeros(max_total, dtype=torch.uint8, device=device)
    if blobs:
        payload[:totals[rank]] = torch.cat(blobs)
    receive = dst is None or rank == dst
    if dst is None:
        bufs = [torch.empty(max_total, dtype=torch.uint8, device=device) for _ in range(world)]
        dist.all_gather(bufs, payload, group=group)
    else:
        bufs = [torch.empty(max_total, dtype=torch.uint8, device=device) for _ in range(world)] if receive else None
        dist.gather(payload, bufs, dst=dst, group=group)
    if not receive:
        return None
    out: List[List[torch.Tensor]] = []
    for r in range(world):
        off, items = 0, []
        for k in range(counts_host[r]):
            sz = int(sizes_host[r][k])
            items.append(bufs[r][off:off + sz])
            off += sz
        out.append(items)
    return out


# ----------------------------------------------------------------------------- driver
_WORKER_STREAMS = {}


def _worker_streams(dev: torch.device, count: int):
    """The job's streams, created once per device: torch recycles a pool of 32 streams per device round
    robin, so making fresh ones on every call would sooner or later alias two workers (or a worker and
    the graph-capture stream) onto one CUDA stream."""
    pool = _WORKER_STREAMS.setdefault(dev.index, [])
    while len(pool) < count:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:count]


def decompose_layers(layers: Sequence[Tuple[str, Callable[[], Tuple[torch.Tensor, Optional[torch.Tensor]]]]],
                     shapes: Sequence[Tuple[int, int]], params, rank: int, world_size: int,
                     device: Optional[torch.device] = None, pack: bool = True, streams: int = 16,
                     **caldera_kwargs):
    """Decomposes this rank's shard of `layers`.

    layers[i] = (name, loader) where loader() returns (W, H) -- generated or loaded directly
    on the owning GPU (or in pinned host memory), so no weight ever crosses ranks.  Returns
    (indices, results) with results[j] the blob (pack=True) or the CalderaDecomposition of
    layer indices[j].  The per-layer seed is derived from the layer index only, so the sharded
    run equals the single-GPU run layer for layer.

    `streams` layers are kept in flight per GPU (one worker thread + CUDA stream each, largest
    layers first): the latency-bound factorisation kernels of one layer overlap with the
    bandwidth- and tensor-bound kernels of the others.  The job always runs in the library's
    "throughput" execution mode (small contraction grids, single-CTA eigensolver), whatever `streams`
    and the world size are, so that a sharded run equals the single-GPU run bit for bit; the previous
    mode is restored on return."""
    import concurrent.futures as cf
    from . import _lib
    from .alg import caldera
    import sys
    previous_mode = _lib.execution_mode()
    previous_switch = sys.getswitchinterval()
    _lib.set_execution_mode("throughput")
    # worker threads alternate between short bursts of Python (a few torch calls) and long GIL-free waits on
    # their stream; with the default 5 ms switch interval a thread that wakes up can sit behind another one's
    # burst for milliseconds while its GPU stream idles
    sys.setswitchinterval(float(os.environ.get("CB_SWITCH_INTERVAL", "0.0002")))
    try:
        return _decompose_layers(layers, shapes, params, rank, world_size, device, pack, streams, caldera, cf,
                                 caldera_kwargs)
    finally:
        sys.setswitchinterval(previous_switch)
        _lib.set_execution_mode(previous_mode)


def _decompose_layers(layers, shapes, params, rank, world_size, device, pack, streams, caldera, cf, caldera_kwargs):
    quantised = params.compute_low_rank_factors and (params.L_bits < 16 or params.R_bits < 16)
    costs = [layer_cost(m, n, params.rank, params.iters, params.lplr_iters, quantised) for (m, n) in shapes]
    mine = lpt_assign(costs, world_size)[rank]
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    nworkers = max(1, min(int(streams), len(mine)))
    order = sorted(range(len(mine)), key=lambda j: (-costs[mine[j]], j))     # big layers first
    results = [None] * len(mine)
    cuda_streams = _worker_streams(dev, nworkers)
    slots = [None] * len(mine)
    if pack and mine:
        # The blobs are the only allocations that outlive a layer.  torch's caching allocator keeps free
        # blocks per stream, so blobs allocated by the workers on their own streams would each be a fresh
        # cudaMalloc (taking the driver's allocation lock while other threads launch graphs: measured 2.5-5 s
        # instead of 1.3 s for the 224-layer job).  One arena, sliced per layer, instead.
        def blob_bytes(m, n):
            q = m * n * max(params.Q_bits, 1) // 8 if params.compute_quantized_component else 0
            quantised_lr = params.L_bits < 16 or params.R_bits < 16
            lr = (m + n) * params.rank * (1 if quantised_lr else 2) if params.compute_low_rank_factors else 0
            return (q + lr + (1 << 15) + 255) // 256 * 256
        sizes = [blob_bytes(*shapes[i]) for i in mine]
        arena = torch.empty(sum(sizes), dtype=torch.uint8, device=dev)
        off = 0
        for j, sz in enumerate(sizes):
            slots[j] = arena[off:off + sz]
            off += sz
    host_time = [[0.0, 0.0] for _ in range(nworkers)]       # seconds inside caldera() / pack per worker

    def work(w):
        try:
            _work(w)
        except BaseException:
            # ThreadPoolExecutor.map re-raises in worker order, not in time order: show every failure as it happens
            import sys
            import traceback
            sys.stderr.write(f"[decompose_layers] worker {w} failed:\n{traceback.format_exc()}\n")
            raise

    def _work(w):
        torch.cuda.set_device(dev)
        with torch.cuda.stream(cuda_streams[w]):
            for j in order[w::nworkers]:
                i = mine[j]
                name, loader = layers[i]
                W, H = loader()
                kw = dict(caldera_kwargs)
                kw.setdefault("seed", 1000 + i)
                kw.setdefault("W_copy", "none")
                kw.setdefault("use_cuda_graph", True)
                if pack:
                    kw.setdefault("return_dense", False)       # the blob holds packed codes and factors only
                t0 = time.perf_counter()
                dec = caldera(params, W, H, device=dev, use_tqdm=False, **kw)
                t1 = time.perf_counter()
                results[j] = pack_decomposition(name, dec, params.Q_bits, params.L_bits, params.R_bits,
                                                tuple(W.shape), out=slots[j]) if pack else dec
                host_time[w][0] += t1 - t0
                host_time[w][1] += time.perf_counter() - t1
            cuda_streams[w].synchronize()

    if nworkers == 1:
        work(0)
    else:
        with cf.ThreadPoolExecutor(max_workers=nworkers) as ex:
            list(ex.map(work, range(nworkers)))
    if os.environ.get("CB_SCHEDULER_TIMES"):
        import sys
        sys.stderr.write(f"[decompose_layers] rank {rank}: per-worker seconds in caldera() "
                         f"{[round(t[0], 3) for t in host_time]}, in pack {[round(t[1], 3) for t in host_time]}\n")
    return mine, results
