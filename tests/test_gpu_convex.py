"""GPU convex_caldera() (cb_convex_prox_iters + post-processing) against the numpy oracle.
The oracle is the documented program, not a run of the reference (which cannot run: no
cvxpy, infeasible cone) -- parity for this entry point is "unpinned" (DESIGN.md section 5)."""
import numpy as np
import pytest
import torch

from oracle import convex_oracle as co
from src.convex_caldera.decomposition.convex_caldera import ConvexCalderaParams, convex_caldera
from ee274_convexcaldera_llm_quantization_b200 import convex_caldera as cc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _problem(seed, m, n, decay=1.0):
    rng = np.random.default_rng(seed)
    U = np.linalg.qr(rng.standard_normal((m, m)))[0]
    V = np.linalg.qr(rng.standard_normal((n, n)))[0]
    k = min(m, n)
    s = 2.0 * np.arange(1, k + 1) ** -decay
    W = ((U[:, :k] * s) @ V[:k, :] + 0.01 * rng.standard_normal((m, n))).astype(np.float32)
    h = (0.5 + rng.random(n)).astype(np.float32)
    return W, h


@pytest.mark.parametrize("m,n,mu,lam,tau,tc", [
    (96, 128, 0.01, 0.5, None, False), (96, 128, 0.1, 0.01, None, False), (128, 96, None, 0.5, 3.0, False),
    (320, 256, 0.02, 1.0, None, True), (320, 256, 0.02, 1.0, None, False)])
def test_matches_oracle(m, n, mu, lam, tau, tc):
    W, h = _problem(m + n, m, n)
    kw = dict(mu=mu, lambda_reg=lam, tau_star=tau, B_tot=4.0, solver_tol=1e-7)
    ref = co.convex_oracle(W, np.diag(h), params=co.ConvexOracleParams(max_iters=1500, **kw))
    rank_cap = min(m, n)                      # no truncation: same program as the dense-SVD oracle
    d = convex_caldera(torch.from_numpy(W), torch.from_numpy(h), params=ConvexCalderaParams(**kw), device=DEV,
                       rank_cap=min(rank_cap, 224), sketch_width=min(rank_cap, 224), power_iters=3, max_iters=1500,
                       check_every=25, use_tensor_cores=tc)
    # bf16 operands make the thresholding inexact at the 2^-8 level: the fixed point of the
    # prox iteration moves by a few 1e-3 in objective value (use_tensor_cores=False is exact to fp32)
    tol = 5e-3 if tc else 5e-4
    assert abs(d.objective_value - ref["objective_value"]) <= tol * abs(ref["objective_value"]), \
        (d.objective_value, ref["objective_value"])
    assert d.b_star[0] == ref["b_star"] and d.b_discrete[0] == ref["b_discrete"]
    assert abs(d.group_info["kappa"] - ref["kappa"]) < 1e-4 * ref["kappa"]
    assert abs(d.group_info["c"] - ref["c"]) < 1e-4 * ref["c"]
    rel_ref = ref["certificates"]["relative_error"]
    assert abs(d.duality_gap - rel_ref) <= max(5e-3, 0.05 * rel_ref), (d.duality_gap, rel_ref)
    # structure of the result (convex_caldera.py:494-514)
    assert d.L_star.shape == d.R_star.shape == d.W_compressed.shape == (m, n)
    assert torch.allclose(d.W_compressed, d.L_star + d.R_star, atol=1e-6)
    L, R = d.group_info["L"], d.group_info["R_lr"]
    assert L.shape == (m, d.effective_rank) and R.shape == (d.effective_rank, n)
    if ref["nuclear_norm"] > 0:
        assert abs(float(d.group_info["singular_values"].sum()) - ref["nuclear_norm"]) <= 2e-2 * ref["nuclear_norm"]
        rel = float((L @ R - d.L_star).norm() / d.L_star.norm().clamp_min(1e-12))
        assert rel < (2e-2 if tc else 2e-3), rel
    else:
        assert d.effective_rank == 0 and float(d.L_star.abs().max()) < 1e-5
    # the residual lives on the reference's grid: integer multiples of delta, clamped
    delta = d.group_info["delta"]
    codes = d.R_star / delta
    assert float((codes - codes.round()).abs().max()) < 1e-3
    assert float(codes.abs().max()) <= 2 ** (int(d.b_discrete[0]) - 1) - 1 + 1e-3


def test_quantize_residual_matches_reference_arithmetic():
    from ee274_convexcaldera_llm_quantization_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(4)
    R = torch.randn(64, 96, generator=g)
    for bits in (2, 3, 4, 8, 16):
        want, delta = co.quantize_residual(R.numpy(), bits)
        Rd = R.to(DEV)
        out = torch.empty_like(Rd)
        dd = torch.empty(2, device=DEV)
        _lib.check(lib.cb_quantize_residual_f32(_lib.ptr(Rd), None, 64, 96, bits, _lib.ptr(out), None, _lib.ptr(dd[0:1]),
                                                _lib.ptr(dd[1:2]), _lib.stream_ptr()), "qres")
        assert abs(float(dd[0]) - delta) <= 1e-7 * delta
        np.testing.assert_array_equal(out.cpu().numpy(), want)


def test_error_conventions():
    W = torch.randn(32, 32)
    with pytest.raises(ValueError):
        convex_caldera(W, params=ConvexCalderaParams(B_tot=1.0, b_min=2.0), device=DEV)
    with pytest.raises(ValueError):
        convex_caldera(W, calibration_data=torch.randn(100, 31), device=DEV)        # wrong feature count
    with pytest.raises(NotImplementedError):
        cc.compute_hessian_and_sensitivities(W, calibration_data=torch.randn(100, 32), device=DEV)   # dense H^(1/2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        convex_caldera(W, device="cpu")


# ------------------------------------------------------------------ pinned on the UNMODIFIED reference
# tests/golden/convex.npz (tests/golden/make_golden_convex.py): the reference's own step functions and its whole
# convex_caldera() through the solver-failure branch (:233-241).
import json  # noqa: E402
import os  # noqa: E402

from ee274_convexcaldera_llm_quantization_b200 import convex_caldera as cc  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "convex.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def test_golden_quantize_residual_bit_exact(gold):
    """quantize_residual (:342-373) on the GPU: delta and every value bit for bit."""
    z, meta = gold
    for c in meta["fn_qres"]:
        Rq, delta = cc.quantize_residual(z[f"fn_qres_in_{c['input']}"], c["bits"], device=DEV)
        assert np.float32(delta) == np.float32(c["delta"]), c
        assert np.array_equal(Rq.cpu().numpy(), z[f"fn_qres_{c['k']}_Rq"]), c


def test_golden_certificates(gold):
    z, meta = gold
    for c in meta["fn_cert"]:
        cert = cc.compute_certificates(z[f"fn_cert_{c['k']}_W"], z[f"fn_cert_{c['k']}_Wc"], 4, 17, 1.25, device=DEV)
        assert cert["avg_bit_width"] == c["avg_bit_width"] and cert["effective_rank"] == c["effective_rank"]
        for key in ("residual_norm", "relative_error", "duality_gap"):
            np.testing.assert_allclose(cert[key], c[key], rtol=2e-6)


def test_golden_calibration(gold):
    """compute_hessian_and_sensitivities (:85-125) for the Hessian kinds this build accepts."""
    z, meta = gold
    for c in meta["fn_calib"]:
        k = c["k"]
        if c["kind"] in ("dense", "calib"):
            continue
        H = torch.from_numpy(z[f"fn_calib_{k}_H"]) if c["kind"] == "diag" else None
        H_sqrt, kappa, cval = cc.compute_hessian_and_sensitivities(torch.from_numpy(z[f"fn_calib_{k}_W"]), H, device=DEV)
        np.testing.assert_allclose(kappa, c["kappa"], rtol=2e-6)
        np.testing.assert_allclose(cval, c["c"], rtol=2e-5)
        g = z[f"fn_calib_{k}_H_sqrt"]
        assert np.abs(H_sqrt.cpu().numpy() - g).max() <= 2e-6 * np.abs(g).max()


def test_golden_low_rank_factorization(gold):
    """low_rank_factorization (:276-339) on the GPU: rank rule exact, L R equal (factors are defined up to the sign
    of each singular pair)."""
    z, meta = gold
    for c in meta["fn_lrf"]:
        L_star = z[f"fn_lrf_in_{c['input']}"]
        Lf, Rf, rank = cc.low_rank_factorization(L_star, c["tau_star"], None if c["tau_star"] is not None else 0.1,
                                                 c["quantize"], c["factor_bits"], device=DEV)
        Lg, Rg = z[f"fn_lrf_{c['k']}_L"], z[f"fn_lrf_{c['k']}_R"]
        assert rank == c["effective_rank"], c
        assert tuple(Lf.shape) == Lg.shape and tuple(Rf.shape) == Rg.shape
        scale = np.abs(Lg @ Rg).max()
        tol = 5e-5 if not c["quantize"] else 3.0 / (2 ** (c["factor_bits"] - 1) - 1)
        assert np.abs((Lf @ Rf).cpu().numpy() - Lg @ Rg).max() <= tol * scale, c
        if not c["quantize"]:
            k = min(rank, 8)
            np.testing.assert_allclose(Lf[:, :k].norm(dim=0).cpu().numpy(), np.linalg.norm(Lg[:, :k], axis=0), rtol=1e-4)


def test_golden_whole_pipeline_fallback_branch(gold):
    """convex_caldera() end to end against the unmodified reference (its solver-failure branch, :233-241)."""
    z, meta = gold
    for c in meta["e2e"]:
        k = c["k"]
        W = torch.from_numpy(z[f"e2e_{k}_W"])
        h = torch.from_numpy(z[f"e2e_{k}_h"]) if c["has_h"] else None
        d = convex_caldera(W, h, params=ConvexCalderaParams(solver="SVD", **c["params"]), device=DEV)
        assert d.solver_status == c["solver_status"] == "failed"
        assert float(d.b_star[0]) == c["b_star"] and int(d.b_discrete[0]) == c["b_discrete"]
        assert d.effective_rank == c["effective_rank"] and d.avg_bit_width == c["avg_bit_width"]
        Lg, Rg = z[f"e2e_{k}_L_star"], z[f"e2e_{k}_R_star"]
        # (fp32 Rayleigh-Ritz through a Gram matrix; the cut at rank 128 falls inside a slowly decaying spectrum)
        assert np.abs(d.L_star.cpu().numpy() - Lg).max() <= 6e-5 * np.abs(Lg).max(), c
        # (atol: when rank 128 >= min(m, n) the "residual" W - L* is pure rounding noise and so is its grid step)
        np.testing.assert_allclose(d.group_info["delta"], c["delta"], rtol=2e-4, atol=1e-5 * float(W.abs().max()))
        # the residual lives on the reference's grid; codes differ only where an SVD rounding difference crosses a
        # rounding boundary (R* = W - L* is the small tail of the spectrum, so those differences are relatively large)
        # R* = W - L* inherits the absolute error of L*
        assert np.abs(d.R_star.cpu().numpy() - Rg).max() <= 6e-5 * np.abs(Lg).max() + c["delta"], c
        # (a 16-bit grid is finer than the fp32 error of L* itself; with min(m, n) <= 128 the fallback keeps every
        #  singular value and the "residual" is rounding noise)
        if c["b_discrete"] <= 8 and min(c["m"], c["n"]) > 128:
            ci, cg = np.rint(d.R_star.cpu().numpy() / d.group_info["delta"]), np.rint(Rg / c["delta"])
            assert np.abs(ci - cg).max() <= 1 and np.mean(ci != cg) <= 0.02, (c, np.mean(ci != cg))
        assert torch.equal(d.W_compressed, d.L_star + d.R_star)
        np.testing.assert_allclose(d.duality_gap, c["duality_gap"], rtol=2e-2, atol=1e-5)
        Lf, Rf = z[f"e2e_{k}_L"], z[f"e2e_{k}_R_lr"]
        L, R = d.group_info["L"], d.group_info["R_lr"]
        assert tuple(L.shape) == Lf.shape and tuple(R.shape) == Rf.shape
        tol = 1e-4 if not c["params"].get("quantize_factors") else 3.0 / (2 ** (c["params"]["factor_bits"] - 1) - 1)
        assert np.abs((L @ R).cpu().numpy() - Lf @ Rf).max() <= tol * np.abs(Lf @ Rf).max()


@pytest.mark.parametrize("mu,tau", [(0.05, None), (None, 0.4)])
def test_gpu_solution_satisfies_kkt(mu, tau):
    """The GPU prox solver's answer against the program itself (not another implementation): one exact float64
    proximal-gradient step from the returned (L*, R*) must leave it unchanged (oracle.kkt_residuals)."""
    rng = np.random.default_rng(5)
    W = (0.02 * rng.standard_normal((72, 96))).astype(np.float32)
    h = (0.5 + rng.random(96)).astype(np.float32)
    kw = dict(mu=mu, tau_star=tau, solver_tol=1e-10)
    d = convex_caldera(torch.from_numpy(W), torch.from_numpy(h), params=ConvexCalderaParams(**kw), device=DEV,
                       rank_cap=72, sketch_width=72, power_iters=3, max_iters=3000, check_every=50, use_tensor_cores=False)
    prm = co.ConvexOracleParams(**kw)
    Hp, lam_max, kappa, c = co.calibrate(W, h)
    rl, rr = co.kkt_residuals(W, d.L_star.cpu().numpy(), d.group_info["R_continuous"].cpu().numpy(), Hp, lam_max, kappa, c, prm)
    assert rl <= 2e-3 and rr <= 2e-3, (rl, rr)


def test_large_shape_satisfies_kkt_and_structure():
    """Config 5 at a size where the whole matrix can still be checked on the host (2048 x 4096, penalty form,
    tensor-core path): the returned point is a fixed point of an exact float64 proximal-gradient step, the low-rank
    part has the planted rank, and the residual lives on the reference's grid."""
    m, n, k = 2048, 4096, 48
    g = torch.Generator().manual_seed(11)
    U = torch.linalg.qr(torch.randn(m, k, generator=g))[0]
    V = torch.linalg.qr(torch.randn(n, k, generator=g))[0]
    s = 24.0 * torch.arange(1, k + 1, dtype=torch.float32) ** -0.3        # 24 ... 7.5
    W = 0.02 * torch.randn(m, n, generator=g) + (U * s) @ V.T
    h = 0.5 + torch.rand(n, generator=g)
    # a direction of size sigma moves from R into L once 2 lambda sigma / kappa > mu (kappa = ||W||_F): put that
    # threshold between the noise (largest singular value ~0.02 (sqrt(m) + sqrt(n)) = 2.2) and the planted part
    kappa = float(W.norm())
    kw = dict(mu=1.0, lambda_reg=1.0 * kappa / (2.0 * 3.0), B_tot=4.0, solver_tol=1e-6)
    d = convex_caldera(W, h, params=ConvexCalderaParams(**kw), device=DEV, rank_cap=96, max_iters=600, check_every=20)
    assert not d.group_info["rank_capped"]
    assert 8 <= d.effective_rank <= k                           # the planted directions above the threshold, no noise
    prm = co.ConvexOracleParams(**kw)
    Hp, lam_max, kappa, c = co.calibrate(W.numpy(), h.numpy())
    rl, rr = co.kkt_residuals(W.numpy(), d.L_star.cpu().numpy(), d.group_info["R_continuous"].cpu().numpy(), Hp, lam_max,
                              kappa, c, prm)
    assert rl <= 2e-2 and rr <= 2e-3, (rl, rr)                  # bf16 contractions inside the thresholding
    codes = d.R_star / d.group_info["delta"]
    assert float((codes - codes.round()).abs().max()) < 1e-3
    assert torch.equal(d.W_compressed, d.L_star + d.R_star)


def _correlated_activations(rng, samples, n, rank):
    """Calibration activations with a strongly non-diagonal Gram matrix (a few shared directions plus noise)."""
    basis = rng.standard_normal((rank, n))
    return (rng.standard_normal((samples, rank)) @ basis / np.sqrt(rank) + 0.3 * rng.standard_normal((samples, n))).astype(np.float32)


@pytest.mark.parametrize("mu,tau,lam", [(0.05, None, 0.5), (None, 0.4, 0.01)])
def test_dense_hessian_solution_satisfies_kkt(mu, tau, lam):
    """SURVEY 8 row C1/C2 with a dense Hessian (convex_caldera.py:103-117): the solver only uses products with H; its
    answer must be a fixed point of an exact float64 proximal-gradient step of the program with that dense H."""
    rng = np.random.default_rng(9)
    m, n = 72, 96
    W = (0.02 * rng.standard_normal((m, n))).astype(np.float32)
    X = _correlated_activations(rng, 400, n, 6)
    H = (X.T.astype(np.float64) @ X.astype(np.float64) / 400).astype(np.float32)
    assert np.abs(H - np.diag(np.diag(H))).max() > 0.2 * np.abs(np.diag(H)).max()        # genuinely dense
    kw = dict(mu=mu, tau_star=tau, lambda_reg=lam, solver_tol=1e-10)     # (lambda_reg: a rank-6 L* in the penalty form)
    d = convex_caldera(torch.from_numpy(W), torch.from_numpy(H), params=ConvexCalderaParams(**kw), device=DEV,
                       rank_cap=72, sketch_width=72, power_iters=3, max_iters=6000, check_every=50, use_tensor_cores=False)
    prm = co.ConvexOracleParams(**kw)
    Hp, lam_max, kappa, c = co.calibrate(W, H)
    L, R = d.L_star.cpu().numpy(), d.group_info["R_continuous"].cpu().numpy()
    rl, rr = co.kkt_residuals(W, L, R, Hp, lam_max, kappa, c, prm)
    assert rl <= 3e-3 and rr <= 3e-3, (rl, rr)
    # the objective it reports is the program's objective at that point
    nuc = float(np.linalg.svd(L.astype(np.float64), compute_uv=False).sum())
    q0 = c * np.exp(-prm.k * min(prm.b_max, prm.B_tot))
    obj = co.objective(W.astype(np.float64), L.astype(np.float64), R.astype(np.float64), Hp, mu if mu is not None else 0.0,
                       tau, prm.lambda_reg, kappa, q0, nuc)
    np.testing.assert_allclose(d.objective_value, obj, rtol=2e-4)
    # and the oracle's own solve of the same program lands on the same objective
    Lo, Ro, _, obj_o, _, (_, so, _), _ = co.solve_prox(W, Hp, lam_max, kappa, c, co.ConvexOracleParams(max_iters=6000, **kw))
    np.testing.assert_allclose(d.objective_value, obj_o, rtol=1e-3)
    assert d.effective_rank == int((so > so[0] * 1e-6).sum()) > 0


def test_calibration_data_equals_its_gram_matrix():
    """convex_caldera.py:103-108: H = X^T X when only calibration data is given."""
    rng = np.random.default_rng(10)
    m, n = 64, 128
    W = torch.from_numpy((0.02 * rng.standard_normal((m, n))).astype(np.float32))
    X = torch.from_numpy(_correlated_activations(rng, 256, n, 4) / 16.0)
    kw = dict(mu=0.05, solver_tol=1e-9)
    args = dict(device=DEV, rank_cap=64, sketch_width=64, power_iters=3, max_iters=2000, check_every=50, use_tensor_cores=False)
    a = convex_caldera(W, None, calibration_data=X, params=ConvexCalderaParams(**kw), **args)
    H = (X.double().T @ X.double()).float()
    b = convex_caldera(W, H, params=ConvexCalderaParams(**kw), **args)
    np.testing.assert_allclose(a.objective_value, b.objective_value, rtol=1e-5)
    assert float((a.L_star - b.L_star).abs().max()) <= 1e-4 * float(b.L_star.abs().max()) + 1e-7
    assert a.effective_rank == b.effective_rank


def test_dense_hessian_with_zero_off_diagonal_takes_the_diagonal_path():
    rng = np.random.default_rng(12)
    W = torch.from_numpy((0.02 * rng.standard_normal((64, 96))).astype(np.float32))
    h = torch.from_numpy((0.5 + rng.random(96)).astype(np.float32))
    kw = dict(mu=0.05, solver_tol=1e-8)
    args = dict(device=DEV, rank_cap=64, sketch_width=64, power_iters=3, max_iters=600, check_every=50, use_tensor_cores=False)
    a = convex_caldera(W, h, params=ConvexCalderaParams(**kw), **args)
    b = convex_caldera(W, torch.diag(h), params=ConvexCalderaParams(**kw), **args)
    assert torch.equal(a.L_star, b.L_star) and a.objective_value == b.objective_value


def test_indefinite_hessian_is_rejected():
    rng = np.random.default_rng(13)
    W = torch.from_numpy((0.02 * rng.standard_normal((32, 64))).astype(np.float32))
    A = rng.standard_normal((64, 64))
    H = torch.from_numpy(((A + A.T) / 2).astype(np.float32))              # eigenvalues of both signs
    with pytest.raises(NotImplementedError):
        convex_caldera(W, H, params=ConvexCalderaParams(), device=DEV)


def test_dense_hessian_large_shape_tensor_core_thresholding():
    """Dense-Hessian solve at 1024 x 2048 with the tcgen05 thresholding path: KKT fixed point in float64."""
    m, n, k = 1024, 2048, 24
    g = torch.Generator().manual_seed(21)
    U = torch.linalg.qr(torch.randn(m, k, generator=g))[0]
    V = torch.linalg.qr(torch.randn(n, k, generator=g))[0]
    s = 16.0 * torch.arange(1, k + 1, dtype=torch.float32) ** -0.3
    W = 0.02 * torch.randn(m, n, generator=g) + (U * s) @ V.T
    B = torch.randn(4, n, generator=g)
    H = torch.eye(n) + (0.03 ** 2) * (B.T @ B)              # dense, eigenvalues 1 ... ~2.9
    kappa = float(W.norm())
    kw = dict(mu=1.0, lambda_reg=1.0 * kappa / (2.0 * 2.0), B_tot=4.0, solver_tol=1e-7)
    d = convex_caldera(W, H, params=ConvexCalderaParams(**kw), device=DEV, rank_cap=64, max_iters=1500, check_every=25)
    prm = co.ConvexOracleParams(**kw)
    Hp, lam_max, kappa, c = co.calibrate(W.numpy(), H.numpy())
    rl, rr = co.kkt_residuals(W.numpy(), d.L_star.cpu().numpy(), d.group_info["R_continuous"].cpu().numpy(), Hp, lam_max,
                              kappa, c, prm)
    assert rl <= 2e-2 and rr <= 3e-3, (rl, rr, d.solver_status, d.group_info["iterations"])
    assert not d.group_info["rank_capped"] and 1 <= d.effective_rank <= k
