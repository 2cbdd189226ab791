"""Host-side logic of the layer-sharded scheduler: LPT partition, blob wire format, and the
end-of-run gather over a world_size-2 gloo group on CPU (the N>1 path; on the GPU box the
same code runs over NCCL)."""
import os
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ee274_convexcaldera_llm_quantization_b200 import scheduler as sch


def llama7b_shapes():
    shapes = []
    for _ in range(32):
        shapes += [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
    return shapes


def test_lpt_balance_llama7b():
    shapes = llama7b_shapes()
    assert len(shapes) == 224
    costs = [sch.layer_cost(m, n, 128, 5) for m, n in shapes]
    for world in (1, 2, 4, 8):
        shards = sch.lpt_assign(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(224))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs) + 1e-6       # imbalance <= one layer
        assert shards == sch.lpt_assign(costs, world)             # deterministic
    assert sch.lpt_assign([], 4) == [[], [], [], []]
    assert sch.lpt_assign([1.0], 3) == [[0], [], []]


def fake_dec(seed, m=8, n=12, r=2, quantised=False):
    g = torch.Generator().manual_seed(seed)
    d = SimpleNamespace()
    d.Q_packed = torch.randint(0, 255, ((m * n + 3) // 4,), generator=g, dtype=torch.uint8)
    d.Q_scale = torch.rand(1, 1, generator=g)
    d.L = torch.randn(m, r, generator=g)
    d.R = torch.randn(r, n, generator=g)
    if quantised:
        d.L_packed = torch.randint(0, 255, ((m * r + 1) // 2,), generator=g, dtype=torch.uint8)
        d.R_packed = torch.randint(0, 255, ((r * n + 1) // 2,), generator=g, dtype=torch.uint8)
        d.L_scale = torch.rand(1, 1, generator=g)
        d.R_scale = torch.rand(1, 1, generator=g)
    d.global_scale = 0.02 + seed
    d.best_step = 3
    d.errors = {"Q": [0.9, 0.8], "LR": [0.85, 0.79]}
    return d


@pytest.mark.parametrize("quantised", [False, True])
def test_blob_roundtrip(quantised):
    d = fake_dec(1, quantised=quantised)
    blob = sch.pack_decomposition("layers.0.q_proj", d, 2, 4 if quantised else 16, 4 if quantised else 16, (8, 12))
    assert blob.dtype == torch.uint8 and blob.numel() % 16 == 0
    out = sch.unpack_decomposition(blob)
    assert out["name"] == "layers.0.q_proj" and out["shape"] == [8, 12] and out["errors"] == d.errors
    assert torch.equal(out["Q_packed"], d.Q_packed) and torch.equal(out["Q_scale"], d.Q_scale)
    if quantised:
        assert "L" not in out and torch.equal(out["L_packed"], d.L_packed) and torch.equal(out["R_scale"], d.R_scale)
    else:
        assert torch.equal(out["L"], d.L.half()) and torch.equal(out["R"], d.R.half())
    with pytest.raises(ValueError):
        sch.unpack_decomposition(torch.zeros(64, dtype=torch.uint8))


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = [(8, 12)] * 5
        costs = [sch.layer_cost(m, n, 2, 1) for m, n in shapes]
        mine = sch.lpt_assign(costs, world)[rank]
        blobs = [sch.pack_decomposition(f"layer{i}", fake_dec(i), 2, 16, 16, shapes[i]) for i in mine]
        got = sch.gather_blobs(blobs, dst=0)
        everyone = sch.gather_blobs(blobs, dst=None)
        ok = True
        if rank == 0:
            names = sorted(sch.unpack_decomposition(b)["name"] for r in got for b in r)
            ok = names == [f"layer{i}" for i in range(5)]
            for r in range(world):
                for b, i in zip(got[r], sch.lpt_assign(costs, world)[r]):
                    u = sch.unpack_decomposition(b)
                    ok = ok and torch.equal(u["Q_packed"], fake_dec(i).Q_packed) and u["name"] == f"layer{i}"
        else:
            ok = got is None
        ok = ok and sum(len(r) for r in everyone) == 5
        # fixed-layout job blobs: sizes are a function of (shape, params), so nothing but the payload is exchanged
        prm = SimpleNamespace(compute_quantized_component=True, compute_low_rank_factors=True, Q_bits=2, L_bits=16,
                              R_bits=16, rank=2, iters=1, lplr_iters=1)
        shapes2 = [(8, 12), (16, 12), (8, 12), (8, 24), (16, 12)]
        shards, sizes, rank_bytes, rank_offs = sch.shard_layout(prm, shapes2, world)
        arena = torch.cat([torch.full((sizes[i],), i + 1, dtype=torch.uint8) for i in shards[rank]]) \
            if shards[rank] else torch.empty(0, dtype=torch.uint8)
        sch.warm_up_gather(torch.device("cpu"), dst=0)
        for dst in (0, 1, None):
            full = sch.gather_arena(arena, rank_bytes, dst=dst)
            if dst is None or rank == dst:
                parts = sch.split_gathered(full, shards, sizes)
                ok = ok and sorted(parts) == list(range(5))
                ok = ok and all(bool((parts[i] == i + 1).all()) and parts[i].numel() == sizes[i] for i in parts)
            else:
                ok = ok and full is None
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gather_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_blob_layout_and_fixed_header():
    """The job's blobs have a layout that is a pure function of (shape, parameters): what lets the gather skip the
    size exchange.  Header region fixed at HEADER_BYTES; sections 16-byte aligned; unpack_decomposition reads it."""
    for lb, rb in ((16, 16), (4, 4), (16, 4)):
        prm = SimpleNamespace(compute_quantized_component=True, compute_low_rank_factors=True, Q_bits=2, L_bits=lb,
                              R_bits=rb, rank=8, iters=2, lplr_iters=2)
        total, secs = sch.blob_layout(prm, (64, 96))
        assert total % 256 == 0 and secs[0][3] == sch.HEADER_BYTES
        names = [s_[0] for s_ in secs]
        if lb < 16 or rb < 16:
            assert names == ["Q_packed", "Q_scale", "L_packed", "R_packed", "L_scale", "R_scale"]
            # a 16-bit factor next to a quantised one travels as 16-bit codes (alg.py:144), 2 bytes per element
            assert dict((s_[0], s_[4]) for s_ in secs)["L_packed"] == 64 * 8 * lb // 8
        else:
            assert names == ["Q_packed", "Q_scale", "L", "R"]
        for _, _, _, off, nb in secs:
            assert off % 16 == 0 and off + nb <= total
        # a blob assembled the way decompose_layers does it
        blob = torch.zeros(total, dtype=torch.uint8)
        meta = {"name": "x", "shape": [64, 96], "q_bits": 2, "l_bits": lb, "r_bits": rb, "global_scale": 0.02, "best_step": 1,
                "errors": {"Q": [0.9], "LR": [0.8]},
                "sections": [{"field": f, "dtype": d, "shape": list(shp), "offset": o - sch.HEADER_BYTES, "nbytes": nb}
                             for f, d, shp, o, nb in secs]}
        blob[:sch.HEADER_BYTES] = torch.frombuffer(bytearray(sch._header_bytes(meta)), dtype=torch.uint8)
        for f, d, shp, o, nb in secs:
            blob[o:o + nb] = (hash(f) % 251)
        out = sch.unpack_decomposition(blob)
        assert out["name"] == "x" and out["errors"] == meta["errors"]
        for f, d, shp, o, nb in secs:
            assert tuple(out[f].shape) == tuple(shp) and str(out[f].dtype) == "torch." + d
    # a very long error trajectory is truncated rather than overflowing the fixed header
    long_meta = {"name": "y", "errors": {"Q": [0.123456789] * 400, "LR": [0.5] * 400}, "sections": []}
    assert len(sch._header_bytes(long_meta)) == sch.HEADER_BYTES


def test_batch_plan_splits_evenly():
    from ee274_convexcaldera_llm_quantization_b200.scheduler import batch_plan
    assert batch_plan(0, 16) == []
    assert batch_plan(16, 16) == [16]
    assert batch_plan(17, 16) == [9, 8]
    assert batch_plan(4, 16) == [4]
    assert batch_plan(64, 16) == [16] * 4
    for count in range(1, 70):
        for batch in (1, 3, 16):
            plan = batch_plan(count, batch)
            assert sum(plan) == count and max(plan) <= batch and max(plan) - min(plan) <= 1
            assert len(plan) == -(-count // batch)
