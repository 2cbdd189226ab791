"""NormalFloat quantisers (nf4 / nf2): the numpy oracle against goldens from the reference (CPU), and the
CUDA kernels against the same goldens (-m gpu)."""
import os

import numpy as np
import pytest
import torch

from oracle import caldera_oracle as orc

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "quantizer_nf.npz"))
NAMES = [str(n) for n in Z["names"]]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_nf_matches_reference(name):
    method = str(Z[f"{name}/method"])
    bits, bs = (int(v) for v in Z[f"{name}/meta"])
    idx, scales, shape = orc.quantize_nf(Z[f"{name}/x"], method, bs)
    assert np.array_equal(idx, Z[f"{name}/idx"]) and np.array_equal(scales, Z[f"{name}/scales"])
    assert np.array_equal(orc.dequantize_nf(idx, scales, shape, method), Z[f"{name}/deq"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_nf_matches_reference_bit_exact(name):
    from src.caldera.utils.quantization import QuantizerFactory
    method = str(Z[f"{name}/method"])
    bits, bs = (int(v) for v in Z[f"{name}/meta"])
    x = torch.from_numpy(Z[f"{name}/x"]).cuda()
    q = QuantizerFactory(method=method, block_size=bs).get_quantizer(bits)
    idx, scales, shape = q.quantize_block(x)
    assert idx.dtype == torch.uint8 and tuple(idx.shape) == tuple(Z[f"{name}/idx"].shape)
    assert np.array_equal(idx.cpu().numpy(), Z[f"{name}/idx"])
    assert np.array_equal(scales.cpu().numpy(), Z[f"{name}/scales"])
    deq = q.dequantize_block(idx, scales, shape)
    assert np.array_equal(deq.cpu().numpy(), Z[f"{name}/deq"])


@pytest.mark.gpu
def test_gpu_nf_large_and_errors():
    from src.caldera.utils.quantization import QuantizerFactory
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1024, 4096, generator=g)
    q = QuantizerFactory(method="nf4", block_size=64).get_quantizer(4)
    idx, scales, shape = q.quantize_block(x.cuda())
    ri, rs, _ = orc.quantize_nf(x.numpy(), "nf4", 64)
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(scales.cpu().numpy(), rs)
    whole = QuantizerFactory(method="nf2", block_size=x.numel()).get_quantizer(2)
    idx2, sc2, _ = whole.quantize_block(x.cuda())
    ri2, rs2, _ = orc.quantize_nf(x.numpy(), "nf2", x.numel())
    assert np.array_equal(idx2.cpu().numpy(), ri2) and np.array_equal(sc2.cpu().numpy(), rs2)
    with pytest.raises(ValueError):
        QuantizerFactory(method="nf4", block_size=64).get_quantizer(2)
    with pytest.raises(NotImplementedError):
        QuantizerFactory(method="nf4", block_size=64).get_quantizer(4).quantize_block(x.cuda(), return_packed=True)
