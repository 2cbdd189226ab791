"""CPU checks of the Convex-CALDERA oracle (oracle/convex_oracle.py): its prox operators against
brute-force minimisation, optimality of the returned point, and the reference's pure-python
post-processing rules (convex_caldera.py:244-273, 342-373)."""
import numpy as np
import pytest

from oracle import convex_oracle as co


def _problem(seed=0, m=48, n=64):
    rng = np.random.default_rng(seed)
    U = np.linalg.qr(rng.standard_normal((m, m)))[0]
    V = np.linalg.qr(rng.standard_normal((n, n)))[0]
    s = 2.0 * np.arange(1, m + 1) ** -1.0
    W = ((U * s) @ V[:m, :] + 0.01 * rng.standard_normal((m, n))).astype(np.float32)
    h = 0.5 + rng.random(n)
    return W, h


def test_radial_shrink_is_the_prox():
    rng = np.random.default_rng(1)
    V = rng.standard_normal((6, 5))
    for tl, kappa, q0 in [(0.3, 2.0, 0.01), (0.3, 2.0, 100.0), (0.05, 30.0, 0.9), (2.0, 1.0, 5.0)]:
        a = co.radial_shrink(V, tl, kappa, q0)
        f = lambda al: 0.5 * np.sum((al * V - V) ** 2) + tl * max(q0, al * al * np.sum(V * V) / kappa)  # noqa: E731
        grid = np.linspace(0, 1.2, 24001)
        best = grid[np.argmin([f(g) for g in grid])]
        assert abs(a - best) < 1e-4, (a, best)


def test_l1_ball_projection():
    s = np.array([3.0, 2.0, 1.0, 0.2])
    out = co.project_l1_nonneg(s, 4.0)
    assert abs(out.sum() - 4.0) < 1e-12 and (out >= 0).all()
    np.testing.assert_allclose(out, [3 - 2 / 3, 2 - 2 / 3, 1 - 2 / 3, 0.0], atol=1e-12)
    np.testing.assert_array_equal(co.project_l1_nonneg(s, 10.0), s)


@pytest.mark.parametrize("mu,lam,tau", [(0.01, 0.5, None), (0.02, 1.0, None), (None, 0.5, 3.0)])
def test_solution_is_a_minimiser(mu, lam, tau):
    W, h = _problem()
    p = co.ConvexOracleParams(mu=mu, lambda_reg=lam, tau_star=tau, B_tot=4.0, solver_tol=1e-9, max_iters=2000)
    H, lam_max, kappa, c = co.calibrate(W, np.diag(h))
    L, R, b_star, obj, status, (U, s, Vt), iters = co.solve_prox(W, H, lam_max, kappa, c, p)
    assert b_star == 4.0 and status == "optimal"
    q0 = c * np.exp(-p.k * b_star)
    nuc = lambda A: np.linalg.svd(A, compute_uv=False).sum()  # noqa: E731
    f = lambda A, B: co.objective(W.astype(np.float64), A, B, H, mu, tau, lam, kappa, q0, nuc(A))  # noqa: E731
    base = f(L, R)
    assert abs(base - obj) < 1e-9 * max(1.0, abs(obj))
    rng = np.random.default_rng(3)
    for _ in range(20):           # convex: no feasible perturbation may decrease the objective
        dL, dR = rng.standard_normal(L.shape) * 1e-3, rng.standard_normal(R.shape) * 1e-3
        A, B = L + dL, R + dR
        if tau is not None and nuc(A) > tau:
            A = A * (tau / nuc(A))
        assert f(A, B) >= base - 1e-7
    if tau is not None:
        assert nuc(L) <= tau * (1 + 1e-9)


def test_postprocessing_rules():
    assert co.round_bits(4.0, [2, 3, 4, 8, 16], 4.0) == 4
    assert co.round_bits(2.6, [2, 3, 4, 8, 16], 2.9) == 2      # nearest is 3 but over budget -> largest feasible
    assert co.round_bits(9.0, [4, 8, 16], 1.0) == 4            # nothing feasible -> smallest
    R = np.array([[1.0, -0.26, 0.1, 0.0]], dtype=np.float32)
    Rq, delta = co.quantize_residual(R, 2)                     # delta = 2/3, levels {-1,0,1}
    assert abs(delta - 2.0 / 3.0) < 1e-7
    np.testing.assert_allclose(Rq, [[2 / 3, 0.0, 0.0, 0.0]], atol=1e-7)
    Rq4, d4 = co.quantize_residual(R, 4)                       # delta = 2/15, clamp at +-7
    np.testing.assert_allclose(Rq4, np.clip(np.rint(R / d4), -7, 7) * d4)
    with pytest.raises(ValueError):
        co.convex_oracle(np.ones((4, 4), np.float32), params=co.ConvexOracleParams(B_tot=1.0, b_min=2.0))


def test_end_to_end_fields():
    W, h = _problem(2)
    o = co.convex_oracle(W, np.diag(h), params=co.ConvexOracleParams(mu=0.01, lambda_reg=0.5, B_tot=4.0))
    assert o["status"] == "optimal" and o["b_discrete"] == 4
    assert o["L"].shape[1] == o["R_lr"].shape[0] == o["certificates"]["effective_rank"]
    np.testing.assert_allclose(o["L"] @ o["R_lr"], o["L_star"], atol=1e-4)
    np.testing.assert_allclose(o["W_compressed"], o["L_star"] + o["R_star"], atol=1e-6)
    assert 0 < o["certificates"]["relative_error"] < 0.2
