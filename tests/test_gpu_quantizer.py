"""GPU parity of the quantiser kernels (csrc/quant.cu) through the reference-facing API.

Bit-exact against (a) golden vectors produced by the reference itself and (b) the numpy
oracle on seeded inputs; size-independent properties at the BASELINE shape (4096 x 4096)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import caldera_oracle as orc
from src.caldera.utils.quantization import QuantizerFactory, pack_codes, unpack_codes

pytestmark = pytest.mark.gpu

_QZ = np.load(os.path.join(GOLDEN, "quantizer.npz"))
_QNAMES = [str(n) for n in _QZ["names"]]
DEV = "cuda"


def _tdiv(x, lv):
    """True IEEE division on the GPU (torch turns tensor / python_scalar into a multiply by the
    reciprocal on CUDA, which is not what the CPU reference computes)."""
    return x / torch.full((), float(lv), device=x.device)


def _quant(x_t, bits, bs, packed=True):
    q = QuantizerFactory(method="uniform", block_size=bs).get_quantizer(bits)
    return q, q.quantize_block(x_t, return_packed=packed)


@pytest.mark.parametrize("name", _QNAMES)
def test_golden_bit_exact(name):
    x = _QZ[f"{str(_QZ[f'{name}/xref'])}/x"]
    bits, bs = (int(v) for v in _QZ[f"{name}/bits_bs"])
    xt = torch.from_numpy(x).to(DEV)
    if name.startswith("transposed"):
        xt = xt.T.contiguous().T          # same values, non-contiguous strides (alg.py:171 quantises L.T)
        assert not xt.is_contiguous()
    q, (codes, scales, shape, packed) = _quant(xt, bits, bs)
    want_codes = _QZ[f"{name}/codes"]
    assert codes.dtype == (torch.int8 if bits <= 8 else torch.int16)
    assert tuple(codes.shape) == want_codes.shape and tuple(scales.shape) == _QZ[f"{name}/scales"].shape
    np.testing.assert_array_equal(codes.cpu().numpy(), want_codes)
    np.testing.assert_array_equal(scales.cpu().numpy(), _QZ[f"{name}/scales"])
    deq = q.dequantize_block(codes, scales, shape)
    assert tuple(deq.shape) == tuple(shape)
    np.testing.assert_array_equal(deq.cpu().numpy(), _QZ[f"{name}/deq"])
    # packed stream: layout equals the oracle's, and decodes back to the same values
    np.testing.assert_array_equal(packed.cpu().numpy(), orc.pack_codes(want_codes, bits))
    deq2 = q.dequantize_block(packed, scales, shape)
    np.testing.assert_array_equal(deq2.cpu().numpy(), _QZ[f"{name}/deq"])
    back = unpack_codes(packed, bits, codes.numel())
    np.testing.assert_array_equal(back.cpu().numpy(), want_codes.reshape(-1))


@pytest.mark.parametrize("bits", [2, 4, 8, 16])
@pytest.mark.parametrize("bs", [16, 32, 64, 128, 256, 48, 0])
def test_oracle_bit_exact_seeded(bits, bs):
    rows, cols = 768, 1536
    g = torch.Generator().manual_seed(100 + bits)
    x = torch.randn(rows, cols, generator=g) * torch.exp(torch.randn(rows, cols, generator=g))
    x[5, :128] = 0.0
    block = rows * cols if bs == 0 else bs
    q, (codes, scales, shape, packed) = _quant(x.to(DEV), bits, block)
    c_ref, s_ref, _ = orc.quantize_uniform(x.numpy(), bits, block)
    np.testing.assert_array_equal(codes.cpu().numpy(), c_ref)
    np.testing.assert_array_equal(scales.cpu().numpy(), s_ref)
    np.testing.assert_array_equal(packed.cpu().numpy(), orc.pack_codes(c_ref, bits))
    deq = q.dequantize_block(codes, scales, shape)
    np.testing.assert_array_equal(deq.cpu().numpy(), orc.dequantize_uniform(c_ref, s_ref, shape, bits))
    np.testing.assert_array_equal(pack_codes(codes, bits).cpu().numpy(), packed.cpu().numpy())


@pytest.mark.parametrize("bits,bs", [(2, 64), (4, 32), (8, 128), (16, 256), (2, 0)])
def test_oracle_bit_exact_large_tensor(bits, bs):
    """A large tensor (several tiles of 1024 elements per resident warp) against the numpy oracle, with zero blocks,
    tiny values and heavy tails; the shared-memory-staged kernel itself is covered by test_pack_only_equals_oracle."""
    rows, cols = 2048, 4096
    g = torch.Generator().manual_seed(500 + bits)
    x = torch.randn(rows, cols, generator=g) * torch.exp(torch.randn(rows, cols, generator=g))
    x[5, :512] = 0.0
    x[9, 3] = 1e-30
    x[2047, 4095] = -77.0
    block = rows * cols if bs == 0 else bs
    q, (codes, scales, shape, packed) = _quant(x.to(DEV), bits, block)
    c_ref, s_ref, _ = orc.quantize_uniform(x.numpy(), bits, block)
    np.testing.assert_array_equal(codes.cpu().numpy(), c_ref)
    np.testing.assert_array_equal(scales.cpu().numpy(), s_ref)
    np.testing.assert_array_equal(packed.cpu().numpy(), orc.pack_codes(c_ref, bits))
    deq = q.dequantize_block(packed, scales, shape)
    np.testing.assert_array_equal(deq.cpu().numpy(), orc.dequantize_uniform(c_ref, s_ref, shape, bits))


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("bs", [32, 64, 128, 256, 0])
@pytest.mark.parametrize("shape", [(96, 1024), (2048, 4096), (34, 1280)])
def test_pack_only_equals_oracle(bits, bs, shape):
    """return_packed="only" (no int8 codes: the lean instantiation of both quantiser kernels, with the branch-free 2- and
    4-bit paths) against the numpy oracle: ties at exactly half a step, all-zero blocks (scale = epsilon), heavy tails,
    a partial last tile (34 x 1280) and the shared-memory-staged kernel (2048 x 4096)."""
    rows, cols = shape
    g = torch.Generator().manual_seed(900 + bits + rows)
    x = torch.randn(rows, cols, generator=g) * torch.exp(torch.randn(rows, cols, generator=g))
    x[3, :256] = 0.0
    x[4, :256] = torch.tensor([1.0, 0.5, -0.5, 0.25, -0.25, 0.75, -1.0, 0.5000001] * 32)   # ties for every bit width
    x[5, 7] = 1e-30
    block = rows * cols if bs == 0 else bs
    q = QuantizerFactory(method="uniform", block_size=block).get_quantizer(bits)
    packed, scales, shp = q.quantize_block(x.to(DEV), return_packed="only")
    c_ref, s_ref, _ = orc.quantize_uniform(x.numpy(), bits, block)
    np.testing.assert_array_equal(scales.cpu().numpy(), s_ref)
    np.testing.assert_array_equal(packed.cpu().numpy(), orc.pack_codes(c_ref, bits))
    np.testing.assert_array_equal(q.dequantize_block(packed, scales, shp).cpu().numpy(),
                                  orc.dequantize_uniform(c_ref, s_ref, shp, bits))


def test_pack_only_tiny_epsilon_takes_the_general_path():
    """Scales below 1e-30 (possible only with a tiny epsilon) are outside the branch-free ternary path's validity range:
    the warp vote must fall back to the IEEE-divide code."""
    g = torch.Generator().manual_seed(77)
    x = torch.randn(64, 1024, generator=g)
    x[0, :128] = 0.0
    x[1, :64] = 3e-36
    for bits in (2, 4):
        q = QuantizerFactory(method="uniform", block_size=64).get_quantizer(bits)
        packed, scales, shp = q.quantize_block(x.to(DEV), epsilon=1e-37, return_packed="only")
        c_ref, s_ref, _ = orc.quantize_uniform(x.numpy(), bits, 64, eps=1e-37)
        np.testing.assert_array_equal(scales.cpu().numpy(), s_ref)
        np.testing.assert_array_equal(packed.cpu().numpy(), orc.pack_codes(c_ref, bits))


@pytest.mark.parametrize("shape,bs", [((37, 53), 37 * 53), ((3, 7), 7), ((1, 20), 4), ((129, 6), 2)])
@pytest.mark.parametrize("bits", [2, 4, 8, 16])
def test_ragged_shapes(shape, bs, bits):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(*shape, generator=g)
    q, (codes, scales, shp, packed) = _quant(x.to(DEV), bits, bs)
    c_ref, s_ref, _ = orc.quantize_uniform(x.numpy(), bits, bs)
    np.testing.assert_array_equal(codes.cpu().numpy(), c_ref)
    np.testing.assert_array_equal(scales.cpu().numpy(), s_ref)
    np.testing.assert_array_equal(packed.cpu().numpy(), orc.pack_codes(c_ref, bits))
    np.testing.assert_array_equal(q.dequantize_block(packed, scales, shp).cpu().numpy(),
                                  orc.dequantize_uniform(c_ref, s_ref, shp, bits))


@pytest.mark.parametrize("bits,bs", [(2, 64), (4, 64), (2, 0), (4, 0), (8, 128)])
def test_full_size_properties(bits, bs):
    """BASELINE shape 4096 x 4096: properties that do not need the CPU oracle."""
    m = n = 4096
    g = torch.Generator(device=DEV).manual_seed(3)
    x = 0.02 * torch.randn(m, n, generator=g, device=DEV)
    block = m * n if bs == 0 else bs
    lv = 2 ** (bits - 1) - 1
    q, (codes, scales, shape, packed) = _quant(x, bits, block)
    # scales are the block abs-max
    amax = x.reshape(-1, block).abs().amax(dim=1, keepdim=True).clamp_min(1e-8)
    assert torch.equal(scales, amax)
    # codes in range, max element maps to +-levels
    assert int(codes.abs().max()) == lv
    # same arithmetic in torch on the GPU (IEEE div/mul/round-half-even)
    want = torch.round((x.reshape(-1, block) / amax) * lv).to(codes.dtype)
    assert torch.equal(codes, want)
    deq = q.dequantize_block(packed, scales, shape)
    assert torch.equal(deq, (_tdiv(want.float(), lv) * amax).reshape(m, n))
    # error bound and idempotence: re-quantising the dequantised tensor reproduces the codes
    assert float((x - deq).abs().max()) <= float(amax.max()) / (2 * lv) * (1 + 1e-5)
    _, (codes2, scales2, _, packed2) = _quant(deq, bits, block)
    assert torch.equal(codes2, codes) and torch.equal(packed2, packed)
    assert torch.equal(unpack_codes(packed, bits, m * n), codes.reshape(-1))


def test_method_errors_on_gpu():
    # (nf4 / nf2: tests/test_nf_quantizer.py; bbint4 / bbint2: tests/test_bbint_quantizer.py)
    q = QuantizerFactory(method="bbint4", block_size=63).get_quantizer(4)
    with pytest.raises(ValueError):
        q.quantize_block(torch.zeros(4, 63, device=DEV))             # a block must pack into whole bytes
