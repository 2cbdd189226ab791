"""Full-size golden for configuration 3 (11008 x 4096, rank 256, Q 2-bit, L/R 4-bit,
lplr_iters 5) from the UNMODIFIED reference on CPU.  Build container only (needs /root/reference); takes
several minutes.  Keeps scalars only (tests/golden/fullsize_c3.json)."""
import json
import os
import sys
import time

import torch

REF = "/root/reference/rank-constrained-regression-main"
sys.path.insert(0, REF)
from src.caldera.utils.quantization import QuantizerFactory  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams  # noqa: E402
from src.caldera.decomposition.alg import caldera  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
M, N = 11008, 4096
g = torch.Generator().manual_seed(1004)
W = 0.02 * torch.randn(M, N, generator=g, dtype=torch.float32)
h = 0.5 + torch.rand(N, generator=g, dtype=torch.float32)
qf = QuantizerFactory(method="uniform", block_size=64)
params = CalderaParams(compute_quantized_component=True, compute_low_rank_factors=True, Q_bits=2, L_bits=4, R_bits=4,
                       rank=256, iters=2, lplr_iters=5, activation_aware_LR=True, update_order=["Q", "LR"],
                       quant_factory_Q=qf, quant_factory_LR=qf, rand_svd=False, sigma_reg=0.0)
torch.manual_seed(42)
t0 = time.time()
dec = caldera(params, W, torch.diag(h), device="cpu", use_tqdm=False, scale_W=True)
dt = time.time() - t0
errs = {k: [float(e) for e in v] for k, v in dec.errors.items()}
flat = [e for pair in zip(errs["Q"], errs["LR"]) for e in pair]
best = min(range(1, len(flat)), key=lambda i: flat[i])
out = {"config": "11008x4096 rank 256 Q2 L/R4 iters 2 lplr 5 order Q,LR aware exact-SVD, synthetic seed 1004",
       "errors": errs, "global_scale": float(dec.global_scale), "best_step": best, "best_error": flat[best],
       "L_scale": float(dec.L_scale.reshape(-1)[0]), "R_scale": float(dec.R_scale.reshape(-1)[0]),
       "torch": torch.__version__, "reference_cpu_seconds": dt, "threads": torch.get_num_threads()}
with open(os.path.join(OUT, "fullsize_c3.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
