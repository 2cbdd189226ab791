"""Full-size golden for the headline configuration (BASELINE.json config 2), from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; ~2-3 minutes of CPU):

    python tests/golden/make_golden_fullsize.py

Synthetic layer 0 of SURVEY.md section 8d (seed 1000: W = 0.02 N(0,1) 4096 x 4096, h = 0.5 + U(0,1)),
H = diag(h) passed as a dense matrix like main.py:165, CalderaParams(Q 2-bit, L/R 16-bit, rank 128,
5 iterations, update_order Q,LR, activation aware, exact SVD), device="cpu".  Only scalars are kept
(tests/golden/fullsize_c2.json): the error trajectory, global_scale, the Q scale of the best iterate
and the index of the best step -- enough to check the B200 path's iterate-0 quantiser and its best error
at the size the benchmark runs.  Valid for the torch build named in the file.
"""
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
M = N = 4096
g = torch.Generator().manual_seed(1000)
W = 0.02 * torch.randn(M, N, generator=g, dtype=torch.float32)
h = 0.5 + torch.rand(N, generator=g, dtype=torch.float32)
qf = QuantizerFactory(method="uniform", block_size=64)
params = CalderaParams(compute_quantized_component=True, compute_low_rank_factors=True, Q_bits=2, L_bits=16, R_bits=16,
                       rank=128, iters=5, lplr_iters=5, activation_aware_LR=True, update_order=["Q", "LR"],
                       quant_factory_Q=qf, quant_factory_LR=qf, rand_svd=False, sigma_reg=0.0)
torch.manual_seed(42)
t0 = time.time()
dec = caldera(params, W, torch.diag(h), device="cpu", use_tqdm=False, scale_W=True)
dt = time.time() - t0
errs = {k: [float(e) for e in v] for k, v in dec.errors.items()}
flat = [e for pair in zip(errs["Q"], errs["LR"]) for e in pair]
best = min(range(1, len(flat)), key=lambda i: flat[i])       # strict arg-min once both components were updated
out = {"config": "4096x4096 rank 128 Q2 L/R16 iters 5 order Q,LR aware exact-SVD, synthetic layer seed 1000",
       "errors": errs, "global_scale": float(dec.global_scale), "Q_scale": float(dec.Q_scale.reshape(-1)[0]),
       "best_step": best, "best_error": flat[best], "torch": torch.__version__, "reference_cpu_seconds": dt,
       "threads": torch.get_num_threads()}
with open(os.path.join(OUT, "fullsize_c2.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
