"""Golden vectors for the bitsandbytes-style quantisers (method="bbint4" / "bbint2") from the UNMODIFIED reference on
CPU.  Build container only (needs /root/reference):  python tests/golden/make_golden_bbint.py -> tests/golden/quantizer_bbint.npz
(The reference appends to outlier_log.csv in the working directory: the script runs inside a temporary directory.)"""
import os
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference/rank-constrained-regression-main"
sys.path.insert(0, REF)
from src.caldera.utils.quantization import QuantizerFactory  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
out, names = {}, []
g = torch.Generator().manual_seed(17)


def with_outliers(x, idx, scale):
    x = x.clone()
    flat = x.view(-1)
    for k, i in enumerate(idx):
        flat[i] = scale * (1 if k % 2 == 0 else -1)
    return x


a = torch.randn(32, 64, generator=g)
b = 0.02 * torch.randn(48, 128, generator=g)
cases = [("bb4_plain", "bbint4", 4, a, 64),
         ("bb4_outliers", "bbint4", 4, with_outliers(a, [5, 64 * 3 + 17, 64 * 3 + 40, 64 * 31 + 63, 700], 60.0), 64),
         ("bb4_small_values", "bbint4", 4, with_outliers(b, [0, 1000, 1001, 6143], 1.5), 128),
         ("bb4_whole", "bbint4", 4, with_outliers(0.02 * torch.randn(40, 96, generator=g), [7, 2000, 3839], 0.9), 3840),
         ("bb4_kat", "bbint4", 4, torch.arange(16, dtype=torch.float32).reshape(2, 8), 8),
         ("bb4_constant", "bbint4", 4, torch.full((2, 32), 0.25), 32),
         ("bb2_plain", "bbint2", 2, torch.randn(16, 96, generator=g), 32),
         ("bb2_outliers", "bbint2", 2, with_outliers(torch.randn(24, 64, generator=g), [3, 64 * 5, 64 * 5 + 1, 1535], -45.0), 64),
         ("bb2_whole", "bbint2", 2, with_outliers(torch.randn(20, 52, generator=g), [11, 500], 30.0), 1040),
         ("bb2_ties", "bbint2", 2, torch.tensor([[0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 1.4999999]]), 8)]
with tempfile.TemporaryDirectory() as tmp:
    os.chdir(tmp)
    for name, method, bits, x, bs in cases:
        q = QuantizerFactory(method=method, block_size=bs).get_quantizer(bits)
        packed, (bmin, scales, ovals, oidx), shape = q.quantize_block(x)
        deq = q.dequantize_block(packed, (bmin, scales, ovals, oidx), shape)
        out[f"{name}/x"] = x.numpy()
        out[f"{name}/packed"] = packed.numpy()
        out[f"{name}/block_min"] = bmin.numpy()
        out[f"{name}/scales"] = scales.numpy()
        out[f"{name}/outlier_values"] = ovals.numpy()
        out[f"{name}/outlier_indices"] = oidx.numpy()
        out[f"{name}/deq"] = deq.numpy()
        out[f"{name}/meta"] = np.array([bits, bs], dtype=np.int64)
        out[f"{name}/method"] = np.array(method)
        names.append(name)
        print(name, "outliers:", ovals.numel())
out["names"] = np.array(names)
out["torch"] = np.array(torch.__version__)
np.savez_compressed(os.path.join(OUT, "quantizer_bbint.npz"), **out)
print("wrote", len(names), "cases")
