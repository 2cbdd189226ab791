"""Pins the oracle's LPLR iteration (oracle/caldera_oracle.py: lplr_refine) on stage-level goldens from the
UNMODIFIED reference (tests/golden/make_golden_lplr.py -> tests/golden/lplr_stage.npz).  CPU only."""
import json
import os

import numpy as np

from oracle import caldera_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "lplr_stage.npz")


def test_oracle_lplr_iteration_matches_reference():
    z = np.load(GOLD)
    meta = json.loads(bytes(z["meta"]).decode())
    for c in meta:
        nm, m, n, r = c["name"], c["m"], c["n"], c["r"]
        if m * n > 200 * 300:
            continue                      # the large tensor-core case is covered on the GPU; keep the CPU suite quick
        res, h = z[f"{nm}_res"], z[f"{nm}_h"]
        H_sqrt = np.diag(np.sqrt(h)) if c["aware"] else np.diag(h)      # alg.py:49-51, 66-68
        prm = orc.OracleParams(L_bits=c["lb"], R_bits=c["rb"], rank=r, lplr_iters=1, activation_aware_LR=c["aware"])
        R_prev = z[f"{nm}_R0"]
        for k in range(c["iters"]):
            L, R, Lc, Rc, Ls, Rs = orc.lplr_refine(res, H_sqrt.astype(np.float32), None, R_prev, prm)
            sc = z[f"{nm}_{k}_scales"]
            np.testing.assert_allclose([float(np.ravel(Ls)[0]), float(np.ravel(Rs)[0])], sc, rtol=2e-5)
            lm = np.mean(np.ravel(Lc) != z[f"{nm}_{k}_L_idxs"])
            rm = np.mean(np.ravel(Rc) != z[f"{nm}_{k}_R_idxs"])
            assert lm <= 2e-3 and rm <= 5e-3, (nm, k, lm, rm)          # LAPACK lstsq here vs there: last-bit differences
            err = float(np.linalg.norm((res - L @ R) @ H_sqrt))
            np.testing.assert_allclose(err, c["errors"][k], rtol=2e-4)
            R_prev = z[f"{nm}_{k}_R_hat"]
