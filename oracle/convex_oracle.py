"""CPU oracle for the Convex-CALDERA entry point.  TEST INFRASTRUCTURE ONLY.

Parity status: UNPINNED.  The reference's `convex_caldera()` cannot run anywhere: it needs
cvxpy + SCS (un-vendored, un-pinned, absent here; RCR/convex_caldera/decomposition/
convex_caldera.py:9, 208-218), has no test or golden output, and its exponential-cone
constraint `ExpCone(q, c, -k*b*c)` (:198) is infeasible for b > 0.  What is restated here is
the *documented* program (README.md:89-93, CONVEX_CALDERA_IMPLEMENTATION.md:31-49,
objective/constraints at convex_caldera.py:168-205) plus the reference's own pure
torch/numpy pre- and post-processing, which *is* followed line by line:

  calibrate              compute_hessian_and_sensitivities    convex_caldera.py:85-125
  solve_prox             solve_convex_optimization (program)  convex_caldera.py:128-241
  round_bits             round_bit_allocations                convex_caldera.py:244-273
  factorize              low_rank_factorization               convex_caldera.py:276-339
  quantize_residual      quantize_residual                    convex_caldera.py:342-373
  certificates           compute_certificates                 convex_caldera.py:376-419
  convex_oracle          convex_caldera                       convex_caldera.py:422-516

Single-group reduction of the program (p = 1): b only enters through q >= c exp(-k b), so the
optimum takes b* = min(b_max, B_tot / p) (infeasible when that is below b_min) and
q = max(q0, ||R||_F^2 / kappa) with q0 = c exp(-k b*).  What remains is

    min_{L,R}  1/2 ||(W - L - R) H^(1/2)||_F^2  +  mu ||L||_*  +  lambda max(q0, ||R||_F^2 / kappa)

(or ||L||_* <= tau* instead of the penalty), solved by accelerated proximal gradient: the
smooth term has gradient -(W - L - R) H in both blocks (Lipschitz constant 2 lambda_max(H)),
the prox of the nuclear norm is singular-value soft-thresholding (projection onto the
nuclear-norm ball in the constrained form) and the prox of the R term is a radial shrink.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

F64 = np.float64


@dataclass
class ConvexOracleParams:
    """Mirror of ConvexCalderaParams (convex_caldera.py:18-54)."""
    B_tot: float = 2.0
    b_min: float = 2.0
    b_max: float = 16.0
    tau_star: Optional[float] = None
    mu: Optional[float] = 0.1
    lambda_reg: float = 0.01
    k: float = 1.0
    discrete_bits: List[int] = field(default_factory=lambda: [2, 3, 4, 8, 16])
    solver_tol: float = 1e-4
    quantize_factors: bool = False
    factor_bits: int = 16
    max_iters: int = 500


def calibrate(W, H=None, calibration_data=None):
    """convex_caldera.py:85-125.  Returns (H_psd, kappa, c)."""
    n = W.shape[1]
    if H is None:
        H = np.eye(n) if calibration_data is None else calibration_data.T @ calibration_data
    H = np.asarray(H, dtype=F64)
    if H.ndim == 1:
        H = np.diag(H)
    H = (H + H.T) / 2
    lam, V = np.linalg.eigh(H)
    lam = np.maximum(lam, 1e-8)                       # :113
    H = (V * lam[None, :]) @ V.T
    kappa = float(np.linalg.norm(W))                  # :120
    c = float(np.var(W.astype(F64), ddof=1)) * 0.1    # torch.var is unbiased (:123)
    return H, float(lam.max()), kappa, c


def project_l1_nonneg(s: np.ndarray, radius: float) -> np.ndarray:
    """Euclidean projection of a non-negative vector onto {x >= 0, sum x <= radius}."""
    if s.sum() <= radius:
        return s
    u = np.sort(s)[::-1]
    css = np.cumsum(u)
    k = np.nonzero(u * np.arange(1, len(u) + 1) > (css - radius))[0][-1]
    theta = (css[k] - radius) / (k + 1.0)
    return np.maximum(s - theta, 0.0)


def radial_shrink(V: np.ndarray, t_lambda: float, kappa: float, q0: float) -> float:
    """alpha such that prox_{t*lambda*max(q0, ||.||^2/kappa)}(V) = alpha V."""
    v = float(np.sum(V * V))
    if v / kappa <= q0:
        return 1.0
    a1 = 1.0 / (1.0 + 2.0 * t_lambda / kappa)
    if a1 * a1 * v / kappa >= q0:
        return a1
    return float(np.sqrt(q0 * kappa / v))


def objective(W, L, R, H, mu, tau_star, lam, kappa, q0, nuc):
    E = W - L - R
    smooth = 0.5 * float(np.sum((E @ H) * E))
    pen = 0.0 if tau_star is not None else mu * nuc
    return smooth + pen + lam * max(q0, float(np.sum(R * R)) / kappa)


def solve_prox(W, H, lam_max, kappa, c, params: ConvexOracleParams, p: float = 1.0):
    """Accelerated proximal gradient on the reduced program.  Returns
    (L_star, R_star, b_star, objective, status, (U, s, Vt) of L_star, iterations)."""
    W = np.asarray(W, dtype=F64)
    b_star = min(params.b_max, params.B_tot / p)
    if b_star < params.b_min:
        return None, None, params.b_min, float("inf"), "infeasible", None, 0
    q0 = c * np.exp(-params.k * b_star)
    t = 1.0 / (2.0 * lam_max)
    L = np.zeros_like(W)
    R = np.zeros_like(W)
    Lp, Rp = L.copy(), R.copy()
    theta = 1.0
    U = s = Vt = None
    prev_obj = np.inf
    status = "max_iters"
    it = 0
    for it in range(1, params.max_iters + 1):
        theta_next = (1.0 + np.sqrt(1.0 + 4.0 * theta * theta)) / 2.0
        beta = (theta - 1.0) / theta_next
        YL = L + beta * (L - Lp)
        YR = R + beta * (R - Rp)
        G = (W - YL - YR) @ H
        VL = YL + t * G
        VR = YR + t * G
        Lp, Rp = L, R
        U, sv, Vt = np.linalg.svd(VL, full_matrices=False)
        if params.tau_star is not None:
            s = project_l1_nonneg(sv, params.tau_star)
        else:
            s = np.maximum(sv - t * params.mu, 0.0)
        L = (U * s[None, :]) @ Vt
        R = radial_shrink(VR, t * params.lambda_reg, kappa, q0) * VR
        theta = theta_next
        obj = objective(W, L, R, H, params.mu, params.tau_star, params.lambda_reg, kappa, q0, float(s.sum()))
        if abs(prev_obj - obj) <= params.solver_tol * max(abs(obj), 1e-30) and it > 5:
            status = "optimal"
            break
        if obj > prev_obj:          # adaptive restart keeps the accelerated scheme monotone enough
            theta = 1.0
        prev_obj = obj
    return L, R, b_star, obj, status, (U, s, Vt), it


def round_bits(b_star: float, discrete_bits, B_tot: float, p: float = 1.0) -> int:
    """convex_caldera.py:244-273."""
    b = min(discrete_bits, key=lambda x: abs(x - b_star))
    if p * b > B_tot:
        valid = [x for x in discrete_bits if p * x <= B_tot]
        b = max(valid) if valid else min(discrete_bits)
    return int(b)


def factorize(s, U, Vt, tau_star, quantize=False, factor_bits=16):
    """convex_caldera.py:276-339 on an SVD that is already at hand.  Returns (L, R, rank)."""
    if tau_star is not None:
        rank = int(np.searchsorted(np.cumsum(s), tau_star) + 1)
        rank = min(rank, len(s))
    else:
        rank = int(np.sum(s > s[0] * 1e-6)) if s[0] > 0 else 0
    rs = np.sqrt(s[:rank])
    Lf = (U[:, :rank] * rs[None, :]).astype(np.float32)
    Rf = (rs[:, None] * Vt[:rank, :]).astype(np.float32)
    if quantize and rank > 0:
        lv = np.float32(2 ** (factor_bits - 1) - 1)
        for A in (Lf, Rf):
            sc = np.max(np.abs(A))
            A[...] = np.round(A / sc * lv) / lv * sc
    return Lf, Rf, rank


def quantize_residual(R_star, b_discrete: int):
    """convex_caldera.py:342-373.  Returns (R_quantised fp32, delta)."""
    R = np.asarray(R_star, dtype=np.float32)
    t = np.float32(np.max(np.abs(R)))
    delta = np.float32(2) * t / np.float32(2 ** b_discrete - 1) if b_discrete < 16 else t / np.float32(2 ** 15)
    mx = np.float32(2 ** (b_discrete - 1) - 1)
    Rint = np.clip(np.rint(R / delta), -mx, mx)
    return (delta * Rint).astype(np.float32), float(delta)


def certificates(W, Wc, b_discrete, rank, obj):
    """convex_caldera.py:376-419."""
    res = float(np.linalg.norm(W - Wc))
    rel = res / float(np.linalg.norm(W))
    return {"avg_bit_width": b_discrete, "effective_rank": rank, "residual_norm": res, "relative_error": rel,
            "duality_gap": rel, "objective_value": obj}


def convex_oracle(W, H=None, calibration_data=None, params: Optional[ConvexOracleParams] = None):
    params = params or ConvexOracleParams()
    W = np.asarray(W, dtype=np.float32)
    Hp, lam_max, kappa, c = calibrate(W, H, calibration_data)
    L, R, b_star, obj, status, svd, iters = solve_prox(W, Hp, lam_max, kappa, c, params)
    if L is None:
        raise ValueError("bit budget infeasible: B_tot / p < b_min")
    b_disc = round_bits(b_star, params.discrete_bits, params.B_tot)
    U, s, Vt = svd
    Lf, Rf, rank = factorize(s, U, Vt, params.tau_star, params.quantize_factors, params.factor_bits)
    Rq, delta = quantize_residual(R, b_disc)
    Wc = L.astype(np.float32) + Rq                     # full L*, not the truncated factors (:484-485)
    cert = certificates(W, Wc, b_disc, rank, obj)
    return {"L_star": L.astype(np.float32), "R_star": Rq, "W_compressed": Wc, "b_star": b_star,
            "b_discrete": b_disc, "L": Lf, "R_lr": Rf, "delta": delta, "certificates": cert,
            "status": status, "objective_value": obj, "iterations": iters, "kappa": kappa, "c": c,
            "nuclear_norm": float(s.sum()), "R_continuous": R.astype(np.float32)}
