"""Drop-in for RCR/convex_caldera/decomposition/convex_caldera.py on libcaldera_b200.

Same entry point, parameter and result dataclasses as the reference (convex_caldera.py:18-82,
422-516).  The CVXPY/SCS conic solve of `solve_convex_optimization` (:128-241) -- which cannot
run at LLM shapes (two dense m x n variables) and whose coded exponential cone (:198) is
infeasible -- is replaced by an accelerated proximal-gradient solve of the documented program
(README.md:89-93) on the GPU: `cb_convex_prox_iters` (csrc/driver.cu).  The reduction and the
CPU restatement it is tested against are in oracle/convex_oracle.py.  Steps 3-6
(round_bit_allocations :244-273, low_rank_factorization :276-339, quantize_residual :342-373,
compute_certificates :376-419) follow the reference arithmetic.

`params.solver = "SVD"` selects what the reference itself does whenever its solver raises (the `except`
branch, :233-241: L* = rank-min(128, .) truncated SVD of W, R* = W - L*, b* = b_min, status "failed"); that
branch is what tests/golden/convex.npz pins end to end against the unmodified reference.

The step functions are also exported under the reference's names (`compute_hessian_and_sensitivities`,
`round_bit_allocations`, `low_rank_factorization`, `quantize_residual`, `compute_certificates`) and run on
the GPU; they take torch tensors (any device) or numpy arrays.

Hessians: identity, diagonal (a dense matrix whose off-diagonal entries are all zero is routed there) and dense
symmetric positive semi-definite matrices, including the Gram matrix X^T X of `calibration_data` (:103-108).  The
dense path never forms H^(1/2) or an eigendecomposition (:111-117): the solver only needs products with H
(`cb_convex_prox_iters_dense`), the eigenvalue clamp at 1e-8 (:113) becomes a diagonal shift by
max(0, 1e-8 - lambda_min) and the step size comes from lambda_max, both from a Lanczos run on the GPU
(`cb_convex_dense_prepare`).  An indefinite H (lambda_min < -1e-4 lambda_max), whose negative eigenvalues the
reference would clamp away, raises NotImplementedError; so does the stand-alone
`compute_hessian_and_sensitivities` for a dense H (it would have to return the dense square root).  There is no
CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import time
import warnings
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .alg import _classify_hessian, _resolve_device
from .params import CalderaDecomposition  # noqa: F401  (the reference imports it too, :15)


@dataclass
class ConvexCalderaParams:
    """Parameters for Convex-CALDERA algorithm (convex_caldera.py:18-54)."""
    B_tot: float = field(default=2.0)
    b_min: float = field(default=2.0)
    b_max: float = field(default=16.0)
    tau_star: Optional[float] = field(default=None)
    mu: Optional[float] = field(default=0.1)
    lambda_reg: float = field(default=0.01)
    k: float = field(default=1.0)
    discrete_bits: List[int] = field(default_factory=lambda: [2, 3, 4, 8, 16])
    solver: str = field(default="SCS")
    solver_verbose: bool = field(default=False)
    solver_tol: float = field(default=1e-4)
    tolerance: float = field(default=0.05)
    apply_qat: bool = field(default=False)
    quantize_factors: bool = field(default=False)
    factor_bits: int = field(default=16)


@dataclass
class ConvexCalderaDecomposition:
    """Results from Convex-CALDERA decomposition (convex_caldera.py:57-82)."""
    L_star: torch.Tensor
    R_star: torch.Tensor
    W_compressed: torch.Tensor
    b_star: np.ndarray
    b_discrete: np.ndarray
    avg_bit_width: float
    effective_rank: float
    duality_gap: float
    residual_norm: float
    solve_time: float
    solver_status: str
    objective_value: float
    group_info: Dict = field(default_factory=dict)


def round_bit_allocations(b_star: float, discrete_bits: List[int], B_tot: float, p: float = 1.0) -> int:
    """Step 3 (convex_caldera.py:244-273)."""
    b_discrete = min(discrete_bits, key=lambda x: abs(x - b_star))
    if p * b_discrete > B_tot:
        valid_bits = [b for b in discrete_bits if p * b <= B_tot]
        b_discrete = max(valid_bits) if valid_bits else min(discrete_bits)
    return b_discrete


def _sum_stats(lib, x: torch.Tensor, y: Optional[torch.Tensor] = None):
    acc = torch.zeros(3, dtype=torch.float64, device=x.device)
    _lib.check(lib.cb_sum_stats(_lib.ptr(x), _lib.ptr(y), x.numel(), _lib.ptr(acc), _lib.stream_ptr()), "sum_stats")
    return acc.tolist()


def _as_device_f32(x, dev: torch.device) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(dev, torch.float32).contiguous()


def compute_hessian_and_sensitivities(W, H=None, calibration_data=None, device: str = "cuda"):
    """Step 1 (convex_caldera.py:85-125): (H_sqrt, kappa, c).  kappa = ||W||_F and c = 0.1 var(W) come from one
    fused reduction on the GPU.  H_sqrt is returned as the reference does (dense n x n) for an identity or
    diagonal Hessian -- the solver itself never forms it (it keeps the diagonal)."""
    dev = _resolve_device(device, W if torch.is_tensor(W) else torch.empty(0))
    lib = _lib.load()
    with torch.cuda.device(dev):
        Wd = _as_device_f32(W, dev)
        n = int(Wd.shape[1])
        if H is None and calibration_data is not None:
            H = _gram_matrix(lib, calibration_data, n, dev)
        h_kind, Hd = _classify_hessian(None if H is None else _as_device_f32(H, dev), n, dev)
        if h_kind == _lib.CB_H_DENSE:
            raise NotImplementedError("compute_hessian_and_sensitivities: the dense square root of a non-diagonal Hessian "
                                      "is not formed on the B200 path (convex_caldera() itself accepts dense Hessians: "
                                      "its solver only needs products with H)")
        hdiag = torch.ones(n, dtype=torch.float32, device=dev) if Hd is None else Hd.clamp_min(1e-8)   # :113
        s1, s2, _ = _sum_stats(lib, Wd)
        numel = Wd.numel()
        kappa = math.sqrt(s2)                                       # torch.norm(W, 'fro') (:120)
        c = 0.1 * (s2 - s1 * s1 / numel) / max(numel - 1, 1)        # torch.var(W) * 0.1, unbiased (:123)
        return torch.diag(hdiag.sqrt()), kappa, c


def _gram_matrix(lib, X, n: int, dev: torch.device) -> torch.Tensor:
    """H = X^T X of calibration activations X (samples x n), convex_caldera.py:108, as one fp32 contraction."""
    Xd = _as_device_f32(X, dev)
    if Xd.dim() != 2 or int(Xd.shape[1]) != n:
        raise ValueError(f"calibration_data must be (samples, {n}), got {tuple(Xd.shape)}")
    N = int(Xd.shape[0])
    H = torch.empty((n, n), dtype=torch.float32, device=dev)
    _lib.check(lib.cb_sgemm_strided(n, n, N, 1.0, _lib.ptr(Xd), 1, n, _lib.ptr(Xd), n, 1, _lib.ptr(H), n, 1, 0,
                                    _lib.stream_ptr()), "sgemm")
    return H


def _prepare_dense_hessian(lib, H: torch.Tensor, n: int, dev: torch.device):
    """(Hs, lambda_min, lambda_max): Hs = (H + H^T)/2 (:111) + max(0, 1e-8 - lambda_min) I (the clamp of :113 as a
    shift), extreme eigenvalues by Lanczos on the GPU.  lambda_min is that of the symmetrised input."""
    Hs = torch.empty((n, n), dtype=torch.float32, device=dev)
    stats = torch.zeros(3, dtype=torch.float32, device=dev)
    nbytes = lib.cb_min_eig_shift_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.cb_convex_dense_prepare(_lib.ptr(H), n, 1e-8, _lib.ptr(Hs), _lib.ptr(stats), _lib.ptr(ws), nbytes,
                                           _lib.stream_ptr()), "convex_dense_prepare")
    shift, lam_min, lam_max = stats.tolist()
    return Hs, lam_min, lam_max + shift


def _truncated_svd(lib, A: torch.Tensor, r: int, q: int, power_iters: int, seed: int):
    """U sqrt(S) (m x r), sqrt(S) V^T (r x n) and S (r) of A by the library's subspace iteration + Rayleigh-Ritz
    (cb_lowrank_init, not activation aware).  q == min(m, n) makes it a full (not randomized) decomposition."""
    m, n = int(A.shape[0]), int(A.shape[1])
    dev = A.device
    Lf = torch.empty((m, r), dtype=torch.float32, device=dev)
    Rf = torch.empty((r, n), dtype=torch.float32, device=dev)
    sig = torch.empty(r, dtype=torch.float32, device=dev)
    nb = int(lib.cb_lowrank_init_workspace_bytes(m, n, r, q, _lib.CB_H_IDENTITY))
    if nb == 0:
        raise ValueError(f"truncated SVD: rank {r} / sketch width {q} not supported for a {m} x {n} matrix (q <= 512)")
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    _lib.check(lib.cb_lowrank_init(_lib.ptr(A), m, n, None, _lib.CB_H_IDENTITY, r, q, int(power_iters), int(seed), 0,
                                   _lib.ptr(Lf), _lib.ptr(Rf), _lib.ptr(sig), _lib.ptr(ws), nb, _lib.stream_ptr()),
               "lowrank_init")
    return Lf, Rf, sig


# Singular values are obtained through the q x q Gram matrix of the projected factor (fp32), which resolves them
# down to ~sqrt(eps_fp32) of the largest one; the reference's "sigma > 1e-6 sigma_1" rank rule (:310-311, applied
# there to a float64 SVD) is therefore applied with this floor.
_SIGMA_FLOOR = 1e-3


def _rank_rule(s_host: np.ndarray, tau_star: Optional[float]) -> int:
    """convex_caldera.py:302-311 on singular values that are already at hand."""
    if len(s_host) == 0 or not s_host[0] > 0:
        return 0
    if tau_star is not None:
        return int(min(np.searchsorted(np.cumsum(s_host), tau_star) + 1, len(s_host)))
    return int(np.sum(s_host > s_host[0] * max(1e-6, _SIGMA_FLOOR)))


def _quantize_factor_(lib, A: torch.Tensor, factor_bits: int) -> None:
    """In place: round(A / max|A| * lv) / lv * max|A| (convex_caldera.py:326-335), one scale per tensor."""
    assert factor_bits in (2, 4, 8, 16), "Bit-width not supported!"
    sc_ = torch.empty(1, dtype=torch.float32, device=A.device)
    _lib.check(lib.cb_quantize_f32(_lib.ptr(A), A.shape[0], A.shape[1], A.stride(0), A.stride(1), factor_bits, 0, 0.0,
                                   None, None, _lib.ptr(sc_), _lib.ptr(A), _lib.stream_ptr()), "quantize factors")


def low_rank_factorization(L_star, tau_star: Optional[float] = None, mu: Optional[float] = None, quantize: bool = False,
                           factor_bits: int = 16, *, device: str = "cuda", rank_cap: int = 512, power_iters: int = 4,
                           seed: int = 0):
    """Step 4 (convex_caldera.py:276-339): L = U sqrt(S), R = sqrt(S) V^T of L*, truncated by the nuclear-norm
    bound (constrained form) or the relative threshold (penalty form), optionally re-quantised.  The SVD runs on
    the GPU (full decomposition when min(m, n) <= rank_cap <= 512, randomized above that).  Returns
    (L, R, effective_rank) like the reference."""
    del mu
    dev = _resolve_device(device, L_star if torch.is_tensor(L_star) else torch.empty(0))
    lib = _lib.load()
    with torch.cuda.device(dev):
        A = _as_device_f32(L_star, dev)
        m, n = int(A.shape[0]), int(A.shape[1])
        r = max(1, min(int(rank_cap), 512, m, n))
        q = r if r == min(m, n) else min(max(2 * r, r + 32), 512, m, n)
        Lf, Rf, sig = _truncated_svd(lib, A, r, q, power_iters, seed)
        rank = _rank_rule(sig.double().cpu().numpy(), tau_star)
        Lf, Rf = Lf[:, :rank].contiguous(), Rf[:rank, :].contiguous()
        if quantize and rank > 0:
            _quantize_factor_(lib, Lf, factor_bits)
            _quantize_factor_(lib, Rf, factor_bits)
    return Lf, Rf, rank


def quantize_residual(R_star, b_discrete: int, *, device: str = "cuda"):
    """Step 5 (convex_caldera.py:342-373): (R_quantized, delta) on the residual grid delta = 2 max|R| / (2^b - 1)."""
    dev = _resolve_device(device, R_star if torch.is_tensor(R_star) else torch.empty(0))
    lib = _lib.load()
    with torch.cuda.device(dev):
        R = _as_device_f32(R_star, dev)
        Rq = torch.empty_like(R)
        dd = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.check(lib.cb_quantize_residual_f32(_lib.ptr(R), None, R.shape[0], R.shape[1], int(b_discrete), _lib.ptr(Rq),
                                                None, _lib.ptr(dd[0:1]), _lib.ptr(dd[1:2]), _lib.stream_ptr()),
                   "quantize_residual")
        return Rq, float(dd[0].item())


def compute_certificates(W, W_compressed, b_discrete: int, effective_rank: float, objective_value: float,
                         p: float = 1.0, *, device: str = "cuda") -> Dict[str, float]:
    """Step 6 (convex_caldera.py:376-419)."""
    del p
    dev = _resolve_device(device, W if torch.is_tensor(W) else torch.empty(0))
    lib = _lib.load()
    with torch.cuda.device(dev):
        _, w_sq, diff_sq = _sum_stats(lib, _as_device_f32(W, dev), _as_device_f32(W_compressed, dev))
    residual_norm = math.sqrt(diff_sq)
    relative_error = residual_norm / math.sqrt(w_sq)
    return {"avg_bit_width": b_discrete, "effective_rank": effective_rank, "residual_norm": residual_norm,
            "relative_error": relative_error, "duality_gap": relative_error, "objective_value": objective_value}


def convex_caldera(
    W: torch.Tensor,
    H: Optional[torch.Tensor] = None,
    calibration_data: Optional[torch.Tensor] = None,
    params: Optional[ConvexCalderaParams] = None,
    device: str = "cuda",
    use_tqdm: bool = False,
    *,
    rank_cap: int = 128,
    sketch_width: int = 0,
    power_iters: int = 2,
    max_iters: int = 300,
    check_every: int = 10,
    seed: int = 0,
    use_tensor_cores: bool = True,
) -> ConvexCalderaDecomposition:
    """Main Convex-CALDERA algorithm (Algorithm 1), convex_caldera.py:422-516.

    Keyword-only extras: rank_cap bounds the rank kept by the randomized singular-value
    thresholding (the reference's dense SVD has no cap; a saturated cap is reported through
    a warning and group_info['rank_capped']); sketch_width / power_iters / seed tune the
    subspace iteration; max_iters / check_every bound the prox loop (the objective is read
    back, i.e. the host synchronises, every `check_every` iterations)."""
    start_time = time.time()
    if params is None:
        params = ConvexCalderaParams()
    if len(W.shape) != 2:
        raise ValueError(f"Support only for 2D matrix, but your input has {len(W.shape)} dimensions.")
    dev = _resolve_device(device, W)
    lib = _lib.load()
    m, n = int(W.shape[0]), int(W.shape[1])
    p = 1.0

    with torch.cuda.device(dev):
        Wd = W.to(dev, torch.float32).contiguous()
        # ---- Step 1: calibration (convex_caldera.py:85-125)
        if H is None and calibration_data is not None:
            H = _gram_matrix(lib, calibration_data, n, dev)              # H = X^T X (:108)
        h_kind, Hd = _classify_hessian(H, n, dev)
        h = None
        lam_max = 1.0
        dense = h_kind == _lib.CB_H_DENSE
        if h_kind == _lib.CB_H_DIAG:
            h = Hd.clamp_min(1e-8).contiguous()          # eigvals = clamp(eigvals, min=1e-8) (:113)
            lam_max = float(h.max().item())
        elif dense:
            h, lam_min, lam_max = _prepare_dense_hessian(lib, Hd, n, dev)
            if lam_min < -1e-4 * lam_max:
                raise NotImplementedError(f"convex_caldera: indefinite Hessian (lambda_min {lam_min:.3e}, lambda_max "
                                          f"{lam_max:.3e}); the reference clamps negative eigenvalues (:113), which "
                                          "needs an eigendecomposition that the B200 path does not form")
            lam_max *= 1.02                              # Ritz values approach lambda_max from below
        s1, s2, _ = _sum_stats(lib, Wd)
        numel = m * n
        kappa = math.sqrt(s2)                             # torch.norm(W, 'fro') (:120)
        c = 0.1 * (s2 - s1 * s1 / numel) / max(numel - 1, 1)   # torch.var(W) * 0.1, unbiased (:123)

        # ---- Step 2: convex solve (reduced program, oracle/convex_oracle.py)
        constrained = params.tau_star is not None
        f32 = dict(dtype=torch.float32, device=dev)
        fallback = str(params.solver).upper() in ("SVD", "FALLBACK")
        if fallback:
            # The `except` branch of solve_convex_optimization (:233-241): what the reference returns whenever
            # its solver raises.  L* = truncated SVD of W with rank min(128, min(m, n)), R* = W - L*.
            r = min(128, m, n)
            q = r if r == min(m, n) else (int(sketch_width) if sketch_width > 0 else min(224, m, n))
            Lf, Rf, sig = _truncated_svd(lib, Wd, r, q, max(int(power_iters), 12), seed)
            L = torch.empty((m, n), **f32)
            R = Wd.clone()
            for dst, alpha, acc in ((L, 1.0, 0), (R, -1.0, 1)):        # L* = (U sqrt S)(sqrt S V^T),  R* = W - L*
                _lib.check(lib.cb_sgemm_strided(m, n, r, alpha, _lib.ptr(Lf), r, 1, _lib.ptr(Rf), n, 1, _lib.ptr(dst), n, 1,
                                                acc, _lib.stream_ptr()), "sgemm")
            b_star, obj, status, done = float(params.b_min), float("inf"), "failed", 0
            svals = torch.cat([sig, torch.ones_like(sig)])        # sigma and the ratios s'/S (nothing was thresholded)
        else:
            b_star = min(params.b_max, params.B_tot / p)
            if b_star < params.b_min:
                raise ValueError(f"convex_caldera: bit budget infeasible (B_tot / p = {params.B_tot / p} < b_min = {params.b_min})")
            mu = -1.0 if constrained else float(params.mu)
            tau = float(params.tau_star) if constrained else 0.0
            q0 = c * math.exp(-params.k * b_star)
            step_t = 1.0 / (2.0 * lam_max)
            r = max(1, min(int(rank_cap), m, n))
            q = int(sketch_width) if sketch_width > 0 else min(max(2 * r, r + 32), m, n)
            if sketch_width <= 0 and q > 224 and r + 32 <= 224:
                q = 224
            q = max(min(q, 512, m, n), r)
            L, Lp, R, Rp = (torch.zeros((m, n), **f32) for _ in range(4))
            Lf = torch.empty((m, r), **f32)
            Rf = torch.empty((r, n), **f32)
            svals = torch.zeros(2 * r, **f32)
            scal = torch.zeros(8, dtype=torch.float64, device=dev)
            ws_bytes = lib.cb_convex_prox_workspace_bytes(m, n, r, q, int(use_tensor_cores))
            if ws_bytes == 0:
                raise ValueError("convex_caldera: invalid rank_cap / sketch_width for this shape")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            theta = C.c_double(1.0)
            prev_obj, obj, status, done, warm = float("inf"), float("inf"), "max_iters", 0, 0
            while done < max_iters:
                step = min(check_every, max_iters - done)
                prox = lib.cb_convex_prox_iters_dense if dense else lib.cb_convex_prox_iters
                st = prox(_lib.ptr(Wd), _lib.ptr(h), m, n, mu, tau, float(params.lambda_reg),
                                              kappa, q0, step_t, r, q, int(power_iters), int(seed) + done, warm,
                                              int(use_tensor_cores), step, C.byref(theta), _lib.ptr(L), _lib.ptr(Lp),
                                              _lib.ptr(R), _lib.ptr(Rp), _lib.ptr(Lf), _lib.ptr(Rf), _lib.ptr(svals),
                                              _lib.ptr(scal), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
                _lib.check(st, "convex_prox_iters")
                warm = 1
                done += step
                nuc, _alpha, r_sq, smooth = scal[:4].tolist()              # host synchronisation
                obj = smooth + (0.0 if constrained else mu * nuc) + params.lambda_reg * max(q0, r_sq / kappa)
                if params.solver_verbose:
                    print(f"[convex_caldera] iter {done}: objective {obj:.6e} nuc {nuc:.4e} ||R||^2 {r_sq:.4e}")
                if abs(prev_obj - obj) <= params.solver_tol * max(abs(obj), 1e-30) and done > check_every:
                    status = "optimal"
                    break
                if obj > prev_obj:
                    theta.value = 1.0                                       # adaptive restart
                prev_obj = obj
            del ws, Lp, Rp

        # ---- Step 3: rounding / repair
        b_discrete = round_bit_allocations(b_star, params.discrete_bits, params.B_tot)

        # ---- Step 4: low-rank factorisation of L* = U diag(s) V^T (convex_caldera.py:276-339)
        sv = svals.cpu()
        s_host, ratio = sv[:r].double().numpy(), svals[r:2 * r]
        if fallback:
            # the reference re-decomposes L* (:298): singular values S[:r] followed by min(m, n) - r zeros
            s_rule = np.concatenate([s_host, np.zeros(min(m, n) - r)])
            rank = min(int(np.searchsorted(np.cumsum(s_rule), params.tau_star) + 1), len(s_rule)) if constrained \
                else int(np.sum(s_host > s_host[0] * 1e-6))
            rank = min(rank, r)       # (columns past r would be the zero singular pairs: they add nothing to L R)
        elif constrained:
            rank = int(min(np.searchsorted(np.cumsum(s_host), params.tau_star) + 1, len(s_host)))
        else:
            rank = int(np.sum(s_host > s_host[0] * 1e-6)) if s_host[0] > 0 else 0
        rank_capped = (not fallback) and bool(s_host[-1] > 0) and r < min(m, n)
        if rank_capped:
            warnings.warn(f"convex_caldera: the thresholded L* saturates rank_cap={r}; increase rank_cap")
        root = ratio.sqrt().contiguous()                   # U sqrt(S) * sqrt(s/S) = U sqrt(s)
        Lfac = torch.empty((m, r), **f32)
        Rfac = torch.empty((r, n), **f32)
        _lib.check(lib.cb_scale_f32(_lib.ptr(Lf), m, r, _lib.ptr(root), 1, 0, _lib.ptr(Lfac), _lib.stream_ptr()), "scale")
        _lib.check(lib.cb_scale_f32(_lib.ptr(Rf), r, n, _lib.ptr(root), 0, 0, _lib.ptr(Rfac), _lib.stream_ptr()), "scale")
        Lfac, Rfac = Lfac[:, :rank].contiguous(), Rfac[:rank, :].contiguous()
        if params.quantize_factors and rank > 0:
            _quantize_factor_(lib, Lfac, params.factor_bits)
            _quantize_factor_(lib, Rfac, params.factor_bits)

        # ---- Step 5: quantise the residual; reconstruction uses the full L* (:484-485)
        R_q = torch.empty((m, n), **f32)
        W_c = torch.empty((m, n), **f32)
        delta_d = torch.empty(2, **f32)
        _lib.check(lib.cb_quantize_residual_f32(_lib.ptr(R), _lib.ptr(L), m, n, int(b_discrete), _lib.ptr(R_q),
                                                _lib.ptr(W_c), _lib.ptr(delta_d[0:1]), _lib.ptr(delta_d[1:2]),
                                                _lib.stream_ptr()), "quantize_residual")

        # ---- Step 6: certificates (convex_caldera.py:376-419)
        _, w_sq, diff_sq = _sum_stats(lib, Wd, W_c)
        delta = float(delta_d[0].item())
    residual_norm = math.sqrt(diff_sq)
    relative_error = residual_norm / math.sqrt(w_sq)
    certificates = {"avg_bit_width": b_discrete, "effective_rank": rank, "residual_norm": residual_norm,
                    "relative_error": relative_error, "duality_gap": relative_error, "objective_value": obj}
    return ConvexCalderaDecomposition(
        L_star=L, R_star=R_q, W_compressed=W_c, b_star=np.array([b_star]), b_discrete=np.array([b_discrete]),
        avg_bit_width=certificates["avg_bit_width"], effective_rank=certificates["effective_rank"],
        duality_gap=certificates["duality_gap"], residual_norm=certificates["residual_norm"],
        solve_time=time.time() - start_time, solver_status=status, objective_value=certificates["objective_value"],
        group_info={"L": Lfac, "R_lr": Rfac, "delta": delta, "certificates": certificates, "iterations": done,
                    "rank_capped": rank_capped, "R_continuous": R, "kappa": kappa, "c": c,
                    "singular_values": sv[:r]})
