"""Golden vectors for the NormalFloat quantisers (method="nf4" / "nf2") from the UNMODIFIED reference on CPU.
Build container only (needs /root/reference):  python tests/golden/make_golden_nf.py -> tests/golden/quantizer_nf.npz"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/rank-constrained-regression-main"
sys.path.insert(0, REF)
from src.caldera.utils.quantization import QuantizerFactory  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
out, names = {}, []
g = torch.Generator().manual_seed(7)
cases = [("nf4_a", "nf4", 4, torch.randn(64, 96, generator=g), 64),
         ("nf4_heavy", "nf4", 4, torch.randn(48, 128, generator=g) * torch.exp(torch.randn(48, 128, generator=g)), 32),
         ("nf4_whole", "nf4", 4, 0.02 * torch.randn(40, 50, generator=g), 2000),
         ("nf4_edges", "nf4", 4, torch.tensor([[1.0, -1.334, 0.056, 0.0560001, -0.056, 0.169, 0.892, -1.167]]), 8),
         ("nf4_zero", "nf4", 4, torch.zeros(2, 16), 16),
         ("nf2_a", "nf2", 2, torch.randn(32, 64, generator=g), 64),
         ("nf2_edges", "nf2", 2, torch.tensor([[0.8165, 0.0, 1e-9, -1e-9, 0.5749, -0.5749, 0.57491, -1.0]]), 8),
         ("nf2_whole", "nf2", 2, torch.randn(24, 20, generator=g), 480)]
for name, method, bits, x, bs in cases:
    q = QuantizerFactory(method=method, block_size=bs).get_quantizer(bits)
    idx, scales, shape = q.quantize_block(x)
    deq = q.dequantize_block(idx, scales, shape)
    out[f"{name}/x"] = x.numpy()
    out[f"{name}/idx"] = idx.numpy()
    out[f"{name}/scales"] = scales.numpy()
    out[f"{name}/deq"] = deq.numpy()
    out[f"{name}/meta"] = np.array([bits, bs], dtype=np.int64)
    out[f"{name}/method"] = np.array(method)
    names.append(name)
out["names"] = np.array(names)
out["torch"] = np.array(torch.__version__)
np.savez_compressed(os.path.join(OUT, "quantizer_nf.npz"), **out)
print("wrote", len(names), "cases")
