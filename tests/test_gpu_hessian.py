"""Hessian accumulation kernel against the fp64 product the reference forms on the CPU (main.py:307-311)."""
import pytest
import torch

from ee274_convexcaldera_llm_quantization_b200.hessian import HessianAccumulator

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("n,Ts", [(512, [256, 64, 8]), (1024, [1000, 333]), (200, [77, 5]), (4096, [2048])])
def test_hessian_accumulate_matches_fp64(n, Ts):
    g = torch.Generator(device=DEV).manual_seed(n + sum(Ts))
    acc = HessianAccumulator(n, DEV)
    ref = torch.zeros(n, n, dtype=torch.float64, device=DEV)
    for T in Ts:
        X = torch.randn(T, n, generator=g, device=DEV) * (1.0 + torch.rand(n, generator=g, device=DEV))
        acc.add(X)
        ref += X.double().T @ X.double()              # a_aT in fp64, summed over samples (main.py:309-313)
    H, h = acc.finalize()
    ref /= len(Ts)                                    # mean over samples (main.py:314)
    scale = float(ref.abs().max())
    assert float((H.double() - ref).abs().max()) < 3e-5 * scale
    assert float((h.double() - ref.diagonal()).abs().max()) < 1e-5 * scale
    assert float((H - H.T).abs().max()) < 1e-5 * scale


def test_hessian_diag_only_feeds_caldera():
    from src.caldera.utils.dataclasses import CalderaParams
    from src.caldera.utils.quantization import QuantizerFactory
    from src.caldera.decomposition.alg import caldera
    g = torch.Generator(device=DEV).manual_seed(3)
    n, m = 384, 256
    acc = HessianAccumulator(n, DEV, dense=False)
    X = torch.randn(4, 100, n, generator=g, device=DEV)
    acc.add(X)
    H, h = acc.finalize(count=400)
    assert H is None
    torch.testing.assert_close(h, (X.reshape(-1, n) ** 2).sum(0) / 400, rtol=1e-5, atol=1e-6)
    W = 0.02 * torch.randn(m, n, generator=g, device=DEV)
    qf = QuantizerFactory(method="uniform", block_size=64)
    p = CalderaParams(Q_bits=2, L_bits=16, R_bits=16, rank=16, iters=2, update_order=["Q", "LR"], quant_factory_Q=qf,
                      quant_factory_LR=qf)
    d = caldera(p, W, h, device=DEV, use_tqdm=False)
    assert d.errors["LR"][0] < d.errors["Q"][0] < 1.0
