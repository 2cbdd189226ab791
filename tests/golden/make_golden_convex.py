"""Goldens for the Convex-CALDERA entry point from the UNMODIFIED reference
(RCR/convex_caldera/decomposition/convex_caldera.py), generated in the build container:

    python tests/golden/make_golden_convex.py        -> tests/golden/convex.npz

The reference module imports cvxpy at line 9 and cvxpy is not installed (and cannot be: no network).
Everything except the conic solve itself is plain torch / numpy, so the module is imported with a
*stand-in* `cvxpy` in sys.modules whose expression objects accept the arithmetic the reference performs
while building its problem (:161-205) and whose `Problem.solve` raises.  That is exactly the situation the
reference handles itself: `solve_convex_optimization` catches the exception and takes its documented
fallback (:233-241: rank-min(128, .) truncated SVD of W, R* = W - L*, b* = b_min, status "failed"), after
which steps 3-6 run unchanged.  So two kinds of goldens come out of the reference's own code:

  fn_*      direct calls of the pure functions with chosen inputs
            compute_hessian_and_sensitivities :85-125, round_bit_allocations :244-273,
            low_rank_factorization :276-339, quantize_residual :342-373, compute_certificates :376-419
  e2e_*     whole `convex_caldera()` runs (:422-516) through the fallback branch

Nothing of the reference is modified or copied; the stand-in only replaces the absent third-party solver.
"""
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = "/root/reference/rank-constrained-regression-main"
OUT = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------ stand-in for the absent cvxpy
class _Expr:
    """Accepts every operation the reference applies while it assembles the problem."""
    __array_ufunc__ = None          # numpy arrays defer to the reflected operators below

    def __init__(self, *a, **k):
        pass

    def _same(self, *a, **k):
        return _Expr()
    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __matmul__ = __rmatmul__ = _same
    __truediv__ = __rtruediv__ = __neg__ = __le__ = __ge__ = __eq__ = __lt__ = __gt__ = _same
    __hash__ = object.__hash__

    @property
    def T(self):
        return _Expr()
    value = None


class SolverError(Exception):
    pass


class _Problem:
    status = None
    value = None

    def __init__(self, objective, constraints=None):
        pass

    def solve(self, *a, **k):
        raise SolverError("cvxpy is not installed in this container (stand-in module)")


def _install_stub():
    cp = types.ModuleType("cvxpy")
    cp.Variable = _Expr
    cp.Minimize = _Expr
    cp.sum_squares = lambda *a, **k: _Expr()
    cp.norm = lambda *a, **k: _Expr()
    cp.maximum = lambda *a, **k: _Expr()
    cp.Problem = _Problem
    cp.SolverError = SolverError
    cp.SCS, cp.MOSEK, cp.ECOS = "SCS", "MOSEK", "ECOS"
    cons = types.ModuleType("cvxpy.constraints")
    cons.ExpCone = _Expr
    cp.constraints = cons
    sys.modules["cvxpy"] = cp
    sys.modules["cvxpy.constraints"] = cons


_install_stub()
sys.path.insert(0, REF)
from src.convex_caldera.decomposition import convex_caldera as ref  # noqa: E402


def decaying_matrix(m, n, seed, alpha=1.0, scale=0.02):
    """Random orthogonal factors, sigma_i ~ i^-alpha: a spectrum a randomized SVD resolves to fp32 accuracy."""
    g = torch.Generator().manual_seed(seed)
    k = min(m, n)
    U, _ = torch.linalg.qr(torch.randn(m, k, generator=g, dtype=torch.float64))
    V, _ = torch.linalg.qr(torch.randn(n, k, generator=g, dtype=torch.float64))
    s = torch.arange(1, k + 1, dtype=torch.float64) ** (-alpha)
    return (scale * (U * s) @ V.T * (k ** 0.5)).float()


def main():
    out = {}
    meta = {"torch": torch.__version__, "numpy": np.__version__, "fn_round": [], "fn_lrf": [], "fn_qres": [],
            "fn_cert": [], "fn_calib": [], "e2e": []}

    # ---- round_bit_allocations (:244-273)
    for b_star in (1.2, 2.0, 2.4, 2.5, 2.6, 3.49, 3.5, 5.9, 6.0, 6.1, 11.9, 12.0, 12.1, 16.0, 40.0):
        for bits in ([2, 3, 4, 8, 16], [4, 8], [16, 2, 8]):
            for B_tot in (1.0, 2.0, 3.0, 4.5, 8.0, 16.0):
                meta["fn_round"].append({"b_star": b_star, "bits": bits, "B_tot": B_tot,
                                         "out": int(ref.round_bit_allocations(b_star, bits, B_tot))})

    # ---- low_rank_factorization (:276-339): exactly low-rank inputs, both rank rules, optional factor quantisation
    k = 0
    for j, (m, n, rk, seed) in enumerate(((96, 128, 12, 1), (160, 96, 24, 2), (128, 160, 40, 3))):
        g = torch.Generator().manual_seed(seed)
        A = torch.randn(m, rk, generator=g, dtype=torch.float64)
        B = torch.randn(rk, n, generator=g, dtype=torch.float64)
        sv = torch.logspace(0, -2, rk, dtype=torch.float64)
        # fp32-representable values (stored once as fp32), handed to the reference as the float64 array cvxpy would return
        L_star = ((A * sv) @ B * 0.01).float().double().numpy()
        out[f"fn_lrf_in_{j}"] = L_star.astype(np.float32)
        S = np.linalg.svd(L_star, compute_uv=False)
        for tau in (None, float(0.6 * S.sum()), float(2.0 * S.sum())):
            for quant, fb in ((False, 16), (True, 8), (True, 4)):
                L, R, er = ref.low_rank_factorization(L_star, tau, None if tau is not None else 0.1, quant, fb)
                out[f"fn_lrf_{k}_L"] = L.numpy()
                out[f"fn_lrf_{k}_R"] = R.numpy()
                meta["fn_lrf"].append({"k": k, "input": j, "tau_star": tau, "quantize": quant, "factor_bits": fb,
                                       "effective_rank": int(er)})
                k += 1

    # ---- quantize_residual (:342-373)
    k = 0
    for j, (m, n, seed) in enumerate(((64, 96, 11), (130, 70, 12))):
        g = torch.Generator().manual_seed(seed)
        Rs = (0.02 * torch.randn(m, n, generator=g)).numpy()
        Rs[0, :8] = np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.5, -3.5], dtype=np.float32) * np.float32(1e-3)   # near ties
        out[f"fn_qres_in_{j}"] = Rs
        Rs = Rs.astype(np.float64)
        for b in (2, 3, 4, 8, 16):
            Rq, delta = ref.quantize_residual(Rs, b)
            out[f"fn_qres_{k}_Rq"] = Rq.numpy()
            meta["fn_qres"].append({"k": k, "input": j, "bits": b, "delta": float(delta)})
            k += 1

    # ---- compute_certificates (:376-419)
    for k, (m, n, seed) in enumerate(((64, 96, 21), (130, 70, 22))):
        g = torch.Generator().manual_seed(seed)
        W = 0.02 * torch.randn(m, n, generator=g)
        Wc = W + 0.003 * torch.randn(m, n, generator=g)
        cert = ref.compute_certificates(W, Wc, 4, 17, 1.25)
        out[f"fn_cert_{k}_W"] = W.numpy()
        out[f"fn_cert_{k}_Wc"] = Wc.numpy()
        meta["fn_cert"].append({"k": k, **{kk: float(v) for kk, v in cert.items()}})

    # ---- compute_hessian_and_sensitivities (:85-125): identity, diagonal, dense, calibration data
    k = 0
    for (m, n, seed, kind) in ((48, 64, 31, "none"), (48, 64, 32, "diag"), (48, 64, 33, "dense"), (48, 64, 34, "calib")):
        g = torch.Generator().manual_seed(seed)
        W = 0.02 * torch.randn(m, n, generator=g)
        H = X = None
        if kind == "diag":
            H = torch.diag(0.5 + torch.rand(n, generator=g))
        elif kind == "dense":
            A = torch.randn(n, n, generator=g)
            H = A @ A.T / n + 0.05 * torch.eye(n)
        elif kind == "calib":
            X = torch.randn(3 * n, n, generator=g)
        H_sqrt, kappa, c = ref.compute_hessian_and_sensitivities(W, H, X)
        out[f"fn_calib_{k}_W"] = W.numpy()
        if H is not None:
            out[f"fn_calib_{k}_H"] = H.numpy()
        if X is not None:
            out[f"fn_calib_{k}_X"] = X.numpy()
        out[f"fn_calib_{k}_H_sqrt"] = H_sqrt.numpy()
        meta["fn_calib"].append({"k": k, "kind": kind, "kappa": float(kappa), "c": float(c)})
        k += 1

    # ---- whole convex_caldera() through the reference's own fallback branch (:233-241)
    cases = [
        dict(m=144, n=160, seed=41, alpha=1.0, params=dict()),                                  # defaults: B_tot 2 -> 2 bits
        dict(m=160, n=144, seed=42, alpha=1.2, params=dict(B_tot=4.0, b_min=3.0)),             # b* = 3 -> 3 bits
        dict(m=136, n=200, seed=43, alpha=0.8, params=dict(B_tot=8.0, b_min=8.0, quantize_factors=True, factor_bits=8)),
        dict(m=152, n=136, seed=44, alpha=1.0, params=dict(B_tot=16.0, b_min=16.0)),           # 16-bit residual grid
        dict(m=144, n=160, seed=45, alpha=1.0, params=dict(tau_star=0.5, mu=None)),            # constrained rank rule
        dict(m=96, n=64, seed=46, alpha=1.0, params=dict(B_tot=4.0, b_min=4.0)),               # min(m, n) < 128
    ]
    for k, c in enumerate(cases):
        W = decaying_matrix(c["m"], c["n"], c["seed"], c["alpha"])
        g = torch.Generator().manual_seed(c["seed"] + 100)
        hdiag = 0.5 + torch.rand(c["n"], generator=g)
        H = torch.diag(hdiag) if k % 2 == 1 else None
        p = ref.ConvexCalderaParams(**c["params"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            d = ref.convex_caldera(W, H, None, p, device="cpu")
        out[f"e2e_{k}_W"] = W.numpy()
        if H is not None:
            out[f"e2e_{k}_h"] = hdiag.numpy()
        out[f"e2e_{k}_L_star"] = d.L_star.numpy()
        out[f"e2e_{k}_R_star"] = d.R_star.numpy()
        assert torch.equal(d.W_compressed, d.L_star + d.R_star)        # :484-485, so it is not stored
        out[f"e2e_{k}_L"] = d.group_info["L"].numpy()
        out[f"e2e_{k}_R_lr"] = d.group_info["R_lr"].numpy()
        meta["e2e"].append({"k": k, "m": c["m"], "n": c["n"], "params": c["params"], "has_h": H is not None,
                            "b_star": float(d.b_star[0]), "b_discrete": int(d.b_discrete[0]),
                            "avg_bit_width": float(d.avg_bit_width), "effective_rank": int(d.effective_rank),
                            "duality_gap": float(d.duality_gap), "residual_norm": float(d.residual_norm),
                            "solver_status": d.solver_status, "delta": float(d.group_info["delta"]),
                            "relative_error": float(d.group_info["certificates"]["relative_error"])})
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "convex.npz"), **out)
    print({k: len(v) if isinstance(v, list) else v for k, v in meta.items()})


if __name__ == "__main__":
    main()
