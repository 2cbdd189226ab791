"""Stage-level goldens for the LPLR loop (update_LR, RCR/caldera/decomposition/alg.py:128-198) from the
UNMODIFIED reference, generated in the build container:

    python tests/golden/make_golden_lplr.py        -> tests/golden/lplr_stage.npz

The reference offers no hook between the iterations of its loop, so this script observes it from outside, without
changing it: the module-level name `quantize_matrix` is wrapped by a recorder (inputs = the unquantised
least-squares solutions L^T and R of alg.py:163 / :175; outputs = codes, scale, dequantised factor), and the
starting factors come from the reference's own LR_init (exact SVD).  For every case the file holds the residual,
the Hessian diagonal, the initial R, and for each of the `lplr_iters` iterations the pre-quantisation L and R,
their codes and scales, the dequantised factors and the inner error ||(res - L R) H_sqrt||_F recomputed from them
with the reference's formula (alg.py:182).  The final state of the reference (best inner iterate, alg.py:184-195)
is stored too.  tests/test_gpu_lplr_stage.py replays each iteration on the GPU from the golden R of the iteration
before, so every iteration is an independent one-stage comparison.
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference/rank-constrained-regression-main"
sys.path.insert(0, REF)
from src.caldera.utils.quantization import QuantizerFactory  # noqa: E402
from src.caldera.utils.dataclasses import CalderaParams, CalderaDecomposition  # noqa: E402
from src.caldera.decomposition import alg  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    dict(name="a4", m=192, n=256, r=16, lb=4, rb=4, aware=True, hkind="uniform", iters=3, seed=1),
    dict(name="a8_4", m=256, n=192, r=24, lb=8, rb=4, aware=True, hkind="heavy", iters=3, seed=2),
    dict(name="n4", m=160, n=224, r=16, lb=4, rb=4, aware=False, hkind="uniform", iters=2, seed=3),
    dict(name="a2", m=128, n=160, r=8, lb=2, rb=2, aware=True, hkind="uniform", iters=2, seed=4),
    dict(name="tc4", m=512, n=768, r=32, lb=4, rb=4, aware=True, hkind="uniform", iters=2, seed=5),   # tensor-core shape
]


def main():
    out, meta = {}, []
    for c in CASES:
        g = torch.Generator().manual_seed(c["seed"])
        m, n, r = c["m"], c["n"], c["r"]
        # a residual with some low-rank structure plus noise, like W - Q in a real layer
        res = 0.02 * torch.randn(m, n, generator=g) + 0.05 * (torch.randn(m, r, generator=g) @ torch.randn(r, n, generator=g)) / r ** 0.5
        h = 0.5 + torch.rand(n, generator=g) if c["hkind"] == "uniform" else torch.exp(1.5 * torch.randn(n, generator=g))
        H = torch.diag(h)
        qf = QuantizerFactory(method="uniform", block_size=64)
        params = CalderaParams(compute_quantized_component=True, compute_low_rank_factors=True, Q_bits=2, L_bits=c["lb"],
                               R_bits=c["rb"], rank=r, iters=1, lplr_iters=c["iters"], activation_aware_LR=c["aware"],
                               update_order=["LR"], quant_factory_Q=qf, quant_factory_LR=qf, rand_svd=False, sigma_reg=0.0)
        # what caldera() prepares before the loop (alg.py:48-68)
        if c["aware"]:
            eigH = torch.linalg.eigh(H)
            H_sqrt = (eigH.eigenvectors @ torch.diag(torch.sqrt(eigH.eigenvalues)) @ eigH.eigenvectors.T)
        else:
            H_sqrt = H
            eigH = torch.return_types.linalg_eigh((torch.ones(n), H))
        info = CalderaDecomposition(Q=torch.zeros(m, n), L=torch.zeros(m, r), R=torch.zeros(r, n))
        info.W = res.clone()
        L0, R0 = alg.LR_init(info, params, H_sqrt, eigH, res)

        calls = []
        real_qm = alg.quantize_matrix

        def recorder(A, quant_params, quant_info):
            o = real_qm(A, quant_params, quant_info)
            calls.append((A.detach().clone(), o.A_idxs.detach().clone(), o.scale.detach().clone(), o.A_hat.detach().clone()))
            return o
        alg.quantize_matrix = recorder
        try:
            alg.update_LR(info, params, res, H_sqrt, eigH, "cpu")
        finally:
            alg.quantize_matrix = real_qm
        assert len(calls) == 2 * c["iters"]
        nm = c["name"]
        out[f"{nm}_res"], out[f"{nm}_h"], out[f"{nm}_R0"] = res.numpy(), h.numpy(), R0.numpy()
        errs = []
        for k in range(c["iters"]):
            LT_pre, L_idx, L_sc, LT_hat = calls[2 * k]
            R_pre, R_idx, R_sc, R_hat = calls[2 * k + 1]
            L_hat = LT_hat.T
            err = float(torch.linalg.matrix_norm((res - L_hat @ R_hat) @ H_sqrt))      # alg.py:182
            errs.append(err)
            out[f"{nm}_{k}_L_pre"] = LT_pre.T.contiguous().numpy()        # m x r
            out[f"{nm}_{k}_L_idxs"] = L_idx.numpy().reshape(-1)            # order of (L^T).flatten()
            out[f"{nm}_{k}_L_hat"] = L_hat.contiguous().numpy()
            out[f"{nm}_{k}_R_pre"] = R_pre.numpy()
            out[f"{nm}_{k}_R_idxs"] = R_idx.numpy().reshape(-1)
            out[f"{nm}_{k}_R_hat"] = R_hat.numpy()
            out[f"{nm}_{k}_scales"] = np.array([float(L_sc.reshape(-1)[0]), float(R_sc.reshape(-1)[0])], dtype=np.float32)
        best = int(np.argmin(errs))   # strict '<' keeps the first minimum (alg.py:184)
        assert torch.equal(info.L_idxs.reshape(-1), calls[2 * best][1].reshape(-1))
        meta.append({**c, "errors": errs, "best": best})
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "lplr_stage.npz"), **out)
    for c in meta:
        print(c["name"], c["errors"], c["best"])


if __name__ == "__main__":
    main()
