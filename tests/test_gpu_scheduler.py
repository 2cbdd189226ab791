"""Layer-sharded job (scheduler.decompose_layers) on the GPU: SURVEY 8(e) determinism check --
the sharded run equals the single-GPU run bit for bit, layer for layer."""
import pytest
import torch

from ee274_convexcaldera_llm_quantization_b200 import CalderaParams, QuantizerFactory, _lib, execution_mode
from ee274_convexcaldera_llm_quantization_b200 import scheduler as sch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _job():
    shapes = [(512, 384), (384, 512), (512, 384), (768, 256), (512, 384), (384, 512)]

    def loader(i):
        def f():
            g = torch.Generator().manual_seed(1000 + i)
            m, n = shapes[i]
            return 0.02 * torch.randn(m, n, generator=g), 0.5 + torch.rand(n, generator=g)
        return f
    layers = [(f"layer{i}", loader(i)) for i in range(len(shapes))]
    qf = QuantizerFactory(method="uniform", block_size=64)
    params = CalderaParams(Q_bits=2, L_bits=4, R_bits=4, rank=16, iters=2, lplr_iters=2, update_order=["Q", "LR"],
                           quant_factory_Q=qf, quant_factory_LR=qf)
    return layers, shapes, params


def test_sharded_job_equals_single_gpu_job_bitwise():
    layers, shapes, params = _job()
    dev = torch.device(DEV, torch.cuda.current_device())
    mode_before = execution_mode()
    idx1, blobs1 = sch.decompose_layers(layers, shapes, params, 0, 1, device=dev, streams=4)
    single = dict(zip(idx1, blobs1))
    assert sorted(single) == list(range(len(shapes)))
    seen = set()
    for rank in range(2):                       # the two ranks of a world-size-2 job, one after the other
        idx, blobs = sch.decompose_layers(layers, shapes, params, rank, 2, device=dev, streams=2)
        for i, b in zip(idx, blobs):
            assert torch.equal(b.cpu(), single[i].cpu()), f"layer {i} differs between sharded and single-GPU run"
            seen.add(i)
    assert seen == set(single)
    assert execution_mode() == mode_before      # the job's throughput mode does not leak
    d = sch.unpack_decomposition(single[3])
    assert tuple(d["shape"]) == (768, 256) and d["name"] == "layer3"


def test_throughput_mode_agrees_with_latency_mode():
    from ee274_convexcaldera_llm_quantization_b200.alg import caldera
    g = torch.Generator().manual_seed(41)
    m, n, r = 2048, 1536, 64
    W = 0.02 * torch.randn(m, n, generator=g)
    h = 0.5 + torch.rand(n, generator=g)
    qf = QuantizerFactory(method="uniform", block_size=64)
    params = CalderaParams(Q_bits=2, L_bits=16, R_bits=16, rank=r, iters=3, update_order=["Q", "LR"],
                           quant_factory_Q=qf, quant_factory_LR=qf)
    a = caldera(params, W, h, device=DEV, use_tqdm=False, seed=3)
    try:
        _lib.set_execution_mode("throughput")
        b = caldera(params, W, h, device=DEV, use_tqdm=False, seed=3)
        c = caldera(params, W, h, device=DEV, use_tqdm=False, seed=3)
    finally:
        _lib.set_execution_mode("latency")
    assert b.errors == c.errors and torch.equal(b.Q_idxs, c.Q_idxs) and torch.equal(b.L, c.L)   # reproducible within a mode
    assert a.errors["Q"][0] == b.errors["Q"][0]                       # no contraction involved yet
    for x, y in zip(a.errors["LR"], b.errors["LR"]):
        assert abs(x - y) <= 2e-3 * abs(x)                            # same algorithm, different summation order
    with pytest.raises(ValueError):
        _lib.set_execution_mode("fast")
