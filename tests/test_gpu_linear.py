"""Packed-decomposition consumer (csrc/gemm_tc.cu: packed_linear_kernel) against the dense reconstruction
the reference uses (main.py:197: W_hat = Q + L @ R)."""
import pytest
import torch

from ee274_convexcaldera_llm_quantization_b200 import _lib
from ee274_convexcaldera_llm_quantization_b200.linear import packed_linear
from src.caldera.utils.dataclasses import CalderaParams
from src.caldera.utils.quantization import QuantizerFactory
from src.caldera.decomposition.alg import caldera

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _codes_case(m, n, bits, T, r, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    lv = 2 ** (bits - 1) - 1
    codes = torch.randint(-lv, lv + 1, (m, n), generator=g, device=DEV, dtype=torch.int32).to(torch.int8)
    lib = _lib.load()
    packed = torch.empty(lib.cb_packed_bytes(m * n, bits), dtype=torch.uint8, device=DEV)
    _lib.check(lib.cb_pack_codes(_lib.ptr(codes), m * n, bits, _lib.ptr(packed), _lib.stream_ptr()), "pack")
    scale = torch.tensor([0.731], device=DEV)
    x = torch.randn(T, n, generator=g, device=DEV)
    L = torch.randn(m, r, generator=g, device=DEV) * 0.1 if r else None
    R = torch.randn(r, n, generator=g, device=DEV) * 0.1 if r else None
    return codes, packed, scale, x, L, R, lv


@pytest.mark.parametrize("m,n,bits,T,r", [(256, 128, 2, 128, 0), (512, 256, 2, 64, 32), (384, 320, 4, 200, 16),
                                          (4096, 4096, 2, 256, 128), (1000, 704, 8, 37, 24), (130, 64, 4, 1, 8)])
def test_packed_linear_matches_dense(m, n, bits, T, r):
    codes, packed, scale, x, L, R, lv = _codes_case(m, n, bits, T, r, seed=m + n + bits + T)
    gs = 0.02
    y = packed_linear(x, packed, scale, bits, m, L, R, global_scale=gs)
    # same operand rounding as the kernel (bf16 x, L, R, t; codes exact), fp64 accumulation
    xb = x.bfloat16().double()
    ref = xb @ codes.double().T * (float(scale) / lv)
    if r:
        t = (xb @ R.bfloat16().double().T).bfloat16().double()
        ref = ref + t @ L.bfloat16().double().T
    ref = ref * gs
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    # r > 0: t = x R^T is rounded to bf16 after an fp32 (kernel) / fp64 (here) accumulation, so a few of its
    # elements land on the other side of a bf16 rounding boundary
    assert err < (5e-4 if r else 2e-5), err
    # and against the plain fp32 dense reconstruction, at bf16 operand accuracy
    Q = codes.float() / lv * scale
    W_hat = (Q + (L @ R if r else 0)) * gs
    dense = x @ W_hat.T
    assert float((y - dense).abs().max() / dense.abs().max()) < 2e-2


def test_packed_linear_from_caldera_result():
    g = torch.Generator().manual_seed(5)
    m, n, r = 512, 384 + 64, 32
    W = 0.02 * torch.randn(m, n, generator=g)
    h = 0.5 + torch.rand(n, generator=g)
    qf = QuantizerFactory(method="uniform", block_size=64)
    p = CalderaParams(Q_bits=2, L_bits=16, R_bits=16, rank=r, iters=2, update_order=["Q", "LR"], quant_factory_Q=qf,
                      quant_factory_LR=qf)
    d = caldera(p, W, h, device=DEV, use_tqdm=False)
    x = torch.randn(3, 50, n, device=DEV)
    y = packed_linear(x, d.Q_packed, d.Q_scale, 2, m, d.L, d.R, global_scale=d.global_scale)
    W_hat = (d.Q + d.L @ d.R) * d.global_scale          # the reference's reconstruction (main.py:197)
    dense = x @ W_hat.T
    assert y.shape == (3, 50, m)
    assert float((y - dense).abs().max() / dense.abs().max()) < 2e-2
    assert float((y - dense).norm() / dense.norm()) < 5e-3
