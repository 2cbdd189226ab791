"""numpy model of the arithmetic the CUDA path performs.  TEST INFRASTRUCTURE ONLY.

Where `caldera_oracle.py` restates the *reference* (eigh of H, full SVD, QR least squares,
four dense products per error), this file restates what `cb_caldera_layer` does on the
device for a diagonal Hessian: column weights instead of H^(1/2), randomized subspace
iteration with CholeskyQR and a Gram Rayleigh-Ritz step instead of the SVD, normal
equations + Cholesky instead of lstsq, and sum_j h_j E_ij^2 instead of the trace.  It lets
the CPU test-suite check that the two formulations agree (tests/test_device_model.py) and
was used to choose the default sketch width / power-iteration counts.
"""
from __future__ import annotations

import numpy as np

from .caldera_oracle import F32, OracleParams, OracleDecomposition, quantize_uniform, dequantize_uniform

import copy


def chol_qr(Z):
    G = Z.T @ Z
    try:
        Lc = np.linalg.cholesky(G.astype(np.float64)).astype(F32)
    except np.linalg.LinAlgError:
        G = G + 1e-6 * np.trace(G) / G.shape[0] * np.eye(G.shape[0], dtype=F32)
        Lc = np.linalg.cholesky(G.astype(np.float64)).astype(F32)
    Linv = np.linalg.inv(Lc.astype(np.float64)).astype(F32)
    return Z @ Linv.T


def subspace_lowrank(Y, r, q, niter, rng, Z0=None):
    """Top-r left singular subspace of Y: returns (Zo[m,q], V[r,q], sig[r], B[q,n])."""
    m, n = Y.shape
    if Z0 is None:
        P = rng.standard_normal((n, q)).astype(F32)
        Z = Y @ P
    else:
        Z = Z0
    Zo = chol_qr(Z)
    for _ in range(niter):
        Po = chol_qr(Y.T @ Zo)
        Zo = chol_qr(Y @ Po)
    Zo = chol_qr(Zo)
    B = Zo.T @ Y
    G = B @ B.T
    lam, V = np.linalg.eigh(G.astype(np.float64))
    order = np.argsort(-lam)[:r]
    return Zo, V[:, order].T.astype(F32), np.sqrt(np.maximum(lam[order], 0)).astype(F32), B


def weighted_err(W, h, Q, L, R, den):
    E = (W - L @ R) - Q
    return float(np.sqrt(np.sum((E * E) * h[None, :], dtype=np.float64) / den))


def caldera_device_model(params: OracleParams, W, h, scale_W=True, q=None, niter=None, seed=0,
                         warm_start=True, global_scale=None, niter_warm=None):
    W = np.asarray(W, dtype=F32)
    m, n = W.shape
    r = params.rank
    if scale_W:
        gs = float(np.sqrt(F32(np.sum(np.square(W.astype(np.float64))) / W.size))) if global_scale is None else global_scale
    else:
        gs = 1
    W = (W / F32(gs)).astype(F32)
    h = np.ones(n, F32) if h is None else np.asarray(h, F32).copy()
    aware = params.activation_aware_LR
    if aware and h.min() < params.sigma_reg:
        h = (h + (F32(params.sigma_reg) - h.min())).astype(F32)
    sh = np.sqrt(h)
    w_inner = h if aware else h * h        # H_sqrt := H when not aware (alg.py:50)
    den = np.sum((W * W) * h[None, :], dtype=np.float64)
    if q is None:
        q = min(max(2 * r, r + 32), min(m, n)) if not params.rand_svd else min(2 * r, min(m, n))
    if niter is None:
        niter = 2 if params.rand_svd else 12       # cold (random start) step, as csrc/driver.cu
        if niter_warm is None:
            niter_warm = 2 if params.rand_svd else 3   # steps warm-started from the previous basis
    rng = np.random.default_rng(seed)

    cur = OracleDecomposition(Q=np.zeros((m, n), F32), L=np.zeros((m, r), F32), R=np.zeros((r, n), F32), W=W)
    best = copy.deepcopy(cur)
    errors = {k: [] for k in params.update_order}
    updated = {k: False for k in params.update_order}
    min_error, step, Zprev = float("inf"), 0, None
    basis_saw_q = False      # the device reuses a basis with the warm count only if it was fitted to W - Q with Q != 0
    for _ in range(params.iters):
        for which in params.update_order:
            if which == "LR" and params.compute_low_rank_factors:
                res = W - cur.Q
                Y = res * sh[None, :] if aware else res
                ni = niter if (Zprev is None or not warm_start or niter_warm is None or not basis_saw_q) else niter_warm
                basis_saw_q = bool(np.any(cur.Q)) or not params.compute_quantized_component
                Zo, V, sig, B = subspace_lowrank(Y, r, q, ni, rng, Zprev if warm_start else None)
                Zprev = Zo       # (the device also rotates it into Ritz vectors; the span is the same)
                if aware:
                    L = Zo @ V.T
                    R = (V @ B) / sh[None, :]
                else:
                    L = (Zo @ V.T) * np.sqrt(sig)[None, :]
                    R = (V @ B) / np.sqrt(sig)[:, None]
                if params.L_bits < 16 or params.R_bits < 16:
                    bestin, best_err = None, float("inf")
                    wl = h if aware else np.ones(n, F32)
                    for _ in range(params.lplr_iters):
                        Rw = R * wl[None, :]
                        Gr = Rw @ R.T
                        Bl = res @ Rw.T
                        L = np.linalg.solve(Gr.astype(np.float64), Bl.T.astype(np.float64)).T.astype(F32)
                        c, s, shp = quantize_uniform(np.ascontiguousarray(L.T), params.L_bits, None)
                        Lc_, Ls_ = c, s
                        L = np.ascontiguousarray(dequantize_uniform(c, s, shp, params.L_bits).T)
                        Gl = L.T @ L
                        Br = L.T @ res
                        R = np.linalg.solve(Gl.astype(np.float64), Br.astype(np.float64)).astype(F32)
                        c, s, shp = quantize_uniform(R, params.R_bits, None)
                        R = dequantize_uniform(c, s, shp, params.R_bits)
                        E = res - L @ R
                        err = np.sqrt(np.sum((E * E) * w_inner[None, :], dtype=np.float64))
                        if err < best_err:
                            best_err, bestin = err, (L, R, Lc_, c, Ls_, s)
                    L, R, cur.L_idxs, cur.R_idxs, cur.L_scale, cur.R_scale = bestin
                cur.L, cur.R = L.astype(F32), R.astype(F32)
            elif which == "Q" and params.compute_quantized_component:
                res = W - cur.L @ cur.R if params.compute_low_rank_factors else W
                c, s, shp = quantize_uniform(res.astype(F32), params.Q_bits, None)
                cur.Q, cur.Q_idxs, cur.Q_scale = dequantize_uniform(c, s, shp, params.Q_bits), c, s
            updated[which] = True
            err = weighted_err(W, h, cur.Q, cur.L, cur.R, den)
            errors[which].append(err)
            if err < min_error and all(updated.values()):
                min_error = err
                best = copy.deepcopy(cur)
                best.best_step = step
            step += 1
    best.errors = errors
    best.global_scale = gs
    return best
