"""Pins oracle/convex_oracle.py (and the host-side pieces of the product's convex_caldera module) on goldens
generated from the UNMODIFIED reference (tests/golden/make_golden_convex.py -> tests/golden/convex.npz).
CPU only; the CUDA path is pinned on the same file in tests/test_gpu_convex.py."""
import json
import os

import numpy as np
import pytest

from oracle import convex_oracle as cvx

GOLD = os.path.join(os.path.dirname(__file__), "golden", "convex.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    return z, json.loads(bytes(z["meta"]).decode())


def test_round_bit_allocations(gold):
    """round_bit_allocations (convex_caldera.py:244-273): oracle and product shim, exact."""
    from ee274_convexcaldera_llm_quantization_b200.convex_caldera import round_bit_allocations
    _, meta = gold
    assert len(meta["fn_round"]) == 270
    for c in meta["fn_round"]:
        assert cvx.round_bits(c["b_star"], c["bits"], c["B_tot"]) == c["out"], c
        assert round_bit_allocations(c["b_star"], c["bits"], c["B_tot"]) == c["out"], c


def test_low_rank_factorization(gold):
    """low_rank_factorization (:276-339): rank rule exact; factors up to the sign of each singular pair, so the
    comparison is on L R and on |L| column by column."""
    z, meta = gold
    for c in meta["fn_lrf"]:
        L_star = z[f"fn_lrf_in_{c['input']}"].astype(np.float64)
        U, s, Vt = np.linalg.svd(L_star, full_matrices=False)
        Lf, Rf, rank = cvx.factorize(s, U, Vt, c["tau_star"], c["quantize"], c["factor_bits"])
        Lg, Rg = z[f"fn_lrf_{c['k']}_L"], z[f"fn_lrf_{c['k']}_R"]
        assert rank == c["effective_rank"] and Lf.shape == Lg.shape and Rf.shape == Rg.shape
        scale = np.abs(Lg @ Rg).max()
        # numerically zero singular pairs (sigma ~ 1e-6 sigma_1 kept by the penalty rule) have arbitrary vectors
        tol = 2e-5 if not c["quantize"] else 3.0 / (2 ** (c["factor_bits"] - 1) - 1)
        assert np.abs(Lf @ Rf - Lg @ Rg).max() <= tol * scale
        if not c["quantize"]:
            k = min(rank, 8)
            np.testing.assert_allclose(np.linalg.norm(Lf[:, :k], axis=0), np.linalg.norm(Lg[:, :k], axis=0), rtol=1e-5)


def test_quantize_residual_bit_exact(gold):
    """quantize_residual (:342-373): same fp32 arithmetic -> same bits."""
    z, meta = gold
    for c in meta["fn_qres"]:
        Rq, delta = cvx.quantize_residual(z[f"fn_qres_in_{c['input']}"].astype(np.float64), c["bits"])
        assert np.float32(delta) == np.float32(c["delta"])
        assert np.array_equal(Rq, z[f"fn_qres_{c['k']}_Rq"]), c


def test_certificates(gold):
    z, meta = gold
    for c in meta["fn_cert"]:
        cert = cvx.certificates(z[f"fn_cert_{c['k']}_W"], z[f"fn_cert_{c['k']}_Wc"], 4, 17, 1.25)
        for key in ("avg_bit_width", "effective_rank", "objective_value"):
            assert cert[key] == c[key]
        for key in ("residual_norm", "relative_error", "duality_gap"):
            np.testing.assert_allclose(cert[key], c[key], rtol=2e-6)


def test_calibration(gold):
    """compute_hessian_and_sensitivities (:85-125): H_sqrt, kappa, c for identity / diagonal / dense / X^T X."""
    z, meta = gold
    for c in meta["fn_calib"]:
        k = c["k"]
        W = z[f"fn_calib_{k}_W"]
        H = z[f"fn_calib_{k}_H"] if f"fn_calib_{k}_H" in z.files else None
        X = z[f"fn_calib_{k}_X"] if f"fn_calib_{k}_X" in z.files else None
        _, _, kappa, cc = cvx.calibrate(W, H, X)
        np.testing.assert_allclose(kappa, c["kappa"], rtol=2e-6)
        np.testing.assert_allclose(cc, c["c"], rtol=2e-5)
        Hs = cvx.hessian_sqrt(W, H, X)
        g = z[f"fn_calib_{k}_H_sqrt"]
        assert np.abs(Hs - g).max() <= 2e-5 * np.abs(g).max()


def test_whole_pipeline_fallback_branch(gold):
    """convex_caldera() end to end through the reference's solver-failure branch (:233-241)."""
    z, meta = gold
    for c in meta["e2e"]:
        k = c["k"]
        W = z[f"e2e_{k}_W"]
        h = z[f"e2e_{k}_h"] if c["has_h"] else None
        prm = {kk: v for kk, v in c["params"].items() if kk in cvx.ConvexOracleParams.__dataclass_fields__}
        o = cvx.convex_oracle(W, h, None, cvx.ConvexOracleParams(**prm), fallback=True)
        assert o["status"] == c["solver_status"] == "failed"
        assert o["b_star"] == c["b_star"] and o["b_discrete"] == c["b_discrete"]
        assert o["certificates"]["effective_rank"] == c["effective_rank"]
        Lg, Rg = z[f"e2e_{k}_L_star"], z[f"e2e_{k}_R_star"]
        assert np.abs(o["L_star"] - Lg).max() <= 1e-5 * np.abs(Lg).max()
        np.testing.assert_allclose(o["delta"], c["delta"], rtol=1e-5)
        # the residual grid: integer codes agree except where an SVD rounding difference crosses a rounding boundary
        ci, cg = np.rint(o["R_star"] / o["delta"]), np.rint(Rg / c["delta"])
        assert np.mean(ci != cg) <= 2e-3 and np.abs(ci - cg).max() <= 1
        np.testing.assert_allclose(o["certificates"]["relative_error"], c["relative_error"], rtol=1e-4)
        np.testing.assert_allclose(o["certificates"]["residual_norm"], c["residual_norm"], rtol=1e-4)
        Lf, Rf = z[f"e2e_{k}_L"], z[f"e2e_{k}_R_lr"]
        assert o["L"].shape == Lf.shape and o["R_lr"].shape == Rf.shape
        tol = 1e-4 if not c["params"].get("quantize_factors") else 3.0 / (2 ** (c["params"]["factor_bits"] - 1) - 1)
        assert np.abs(o["L"] @ o["R_lr"] - Lf @ Rf).max() <= tol * np.abs(Lf @ Rf).max()


def test_prox_solution_satisfies_kkt():
    """The prox solver's answer is checked against the program itself, not against another implementation:
    at the optimum one proximal-gradient step is a fixed point."""
    rng = np.random.default_rng(5)
    W = (0.02 * rng.standard_normal((72, 96))).astype(np.float32)
    h = 0.5 + rng.random(96)
    for prm in (cvx.ConvexOracleParams(mu=0.05, max_iters=3000, solver_tol=1e-13),
                cvx.ConvexOracleParams(mu=None, tau_star=0.4, max_iters=3000, solver_tol=1e-13)):
        Hp, lam_max, kappa, c = cvx.calibrate(W, h)
        L, R, _, _, _, _, _ = cvx.solve_prox(W, Hp, lam_max, kappa, c, prm)
        rl, rr = cvx.kkt_residuals(W, L, R, Hp, lam_max, kappa, c, prm)
        assert rl <= 1e-5 and rr <= 1e-5, (rl, rr)
