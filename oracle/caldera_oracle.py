"""CPU oracle for the CALDERA decomposition hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithm.  It exists so that
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs can check and
time the B200 path against the reference's arithmetic on a box where `/root/reference`
does not exist.  Nothing under `ee274_convexcaldera_llm_quantization_b200/` may import it.

Parity status: PINNED.  `tests/golden/*.npz` were produced by importing the reference
itself (`tests/golden/make_golden.py`, torch 2.11.0 CPU) and `tests/test_oracle_golden.py`
checks every function here against them: bit-exact for codes/scales/dequantised values,
1e-4 relative for the floating-point error trajectories (LAPACK driver differences only).

Citations use RCR/ = /root/reference/rank-constrained-regression-main/src/.

  quantize_uniform      RCR/caldera/utils/quantization.py:244-268, 93-101
  dequantize_uniform    RCR/caldera/utils/quantization.py:290-295, 103-105, 306-307
  pack_codes/unpack     new format; bit order follows the only packing convention in the
                        reference, RCR/caldera/utils/quantization.py:152, 217-220
  hessian_factors       RCR/caldera/decomposition/alg.py:11-23, 44-68
  weighted_error        RCR/caldera/decomposition/alg.py:286-302
  lowrank_init          RCR/caldera/decomposition/alg.py:201-235
  lplr_refine           RCR/caldera/decomposition/alg.py:144-198
  caldera_oracle        RCR/caldera/decomposition/alg.py:24-112
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- quantiser
def uniform_levels(bits: int) -> int:
    """Number of positive levels of the symmetric grid: 2^(b-1) - 1 (quantization.py:95)."""
    return (1 << (bits - 1)) - 1


def quantize_uniform(x: np.ndarray, bits: int, block_size: Optional[int] = None,
                     eps: float = 1e-8):
    """Block-wise abs-max quantiser.  Returns (codes[nblk, bs], scales[nblk, 1], shape).

    Each block of `block_size` consecutive row-major elements gets scale
    s = max(absmax, eps) and codes rint((x / s) * levels) (round-half-even), stored as
    int8 (bits <= 8) or int16.  `block_size=None` means one block for the whole tensor,
    which is what `quantize_matrix` forces (alg.py:247).
    """
    if x.ndim != 2:
        raise ValueError(f"Support only for 2D matrix, but your input has {x.ndim} dimensions.")
    if bits not in (2, 4, 8, 16):
        raise AssertionError("Bit-width not supported!")
    numel = x.shape[0] * x.shape[1]
    bs = numel if block_size is None else int(block_size)
    if numel % bs != 0:
        raise ValueError(f"Weight with shape {x.shape[0]} x {x.shape[1]} is not divisible by block size {bs}")
    blocks = np.ascontiguousarray(x, dtype=F32).reshape(-1, bs)
    amax = np.abs(blocks).max(axis=1, keepdims=True)
    scales = np.maximum(amax, F32(eps)).astype(F32)
    scaled = (blocks / scales) * F32(uniform_levels(bits))
    codes = np.rint(scaled).astype(np.int8 if bits <= 8 else np.int16)
    return codes, scales, tuple(x.shape)


NF_LEVELS = {   # quantization.py:44-50 (nf4), :57-60 (nf2)
    "nf4": np.array([-1.334, -1.0, -0.784, -0.617, -0.476, -0.347, -0.226, -0.112, 0.0, 0.112, 0.226, 0.347, 0.476,
                     0.617, 0.784, 1.0], dtype=F32),
    "nf2": np.array([-0.8165, -0.3333, 0.3333, 0.8165], dtype=F32),
}


def quantize_nf(x: np.ndarray, method: str, block_size: int, eps: float = 1e-8):
    """NormalFloat codebook quantiser (quantization.py:270-279 with _quantize_nf :68-88): per block
    s = max(absmax, eps); index = number of thresholds (midpoints of neighbouring levels, fp32) that x / s
    exceeds.  Returns (indices uint8 [nblk, bs], scales fp32 [nblk, 1], shape)."""
    levels = NF_LEVELS[method]
    thr = ((levels[:-1] + levels[1:]) / F32(2)).astype(F32)
    blocks = np.ascontiguousarray(x, dtype=F32).reshape(-1, int(block_size))
    scales = np.maximum(np.abs(blocks).max(axis=1, keepdims=True), F32(eps)).astype(F32)
    scaled = (blocks / scales).astype(F32)
    idx = np.zeros(blocks.shape, dtype=np.uint8)
    for t in thr:
        idx += (scaled > t).astype(np.uint8)
    return idx, scales, tuple(x.shape)


def dequantize_nf(idx: np.ndarray, scales: np.ndarray, shape, method: str) -> np.ndarray:
    """quantization.py:296-298 with _dequantize_nf (:90-94)."""
    return (NF_LEVELS[method][idx.astype(np.int64)] * scales).astype(F32).reshape(shape)


def quantize_bbint(x: np.ndarray, method: str, block_size: int, eps: float = 1e-8):
    """bitsandbytes-style asymmetric block quantiser (quantization.py:107-154 bbint4, :175-221 bbint2).  Per block:
    outliers |x - mean| > 6 max(std_unbiased, eps) go to a side table (values, (row, col) in torch.nonzero order) and
    are replaced by the mean; scale = max((max - min) / levels, eps); code = clip(rint((x - min) / scale), 0, levels),
    packed MSB-first.  mean / std are accumulated in float64 and rounded to fp32 (torch accumulates in fp32 in an
    order of its own: the two can differ in the last bit, which only matters for an element exactly on the outlier
    threshold).  Returns (packed uint8 [nblk, bs*bits/8], (block_min, scales, outlier_values, outlier_indices), shape)."""
    bits = 4 if method == "bbint4" else 2
    levels = F32(2 ** bits - 1)
    blocks = np.ascontiguousarray(x, dtype=F32).reshape(-1, int(block_size)).copy()
    b64 = blocks.astype(np.float64)
    mean = b64.mean(axis=1, keepdims=True).astype(F32)
    std = np.sqrt(((b64 - mean.astype(np.float64)) ** 2).sum(axis=1, keepdims=True) / (blocks.shape[1] - 1)).astype(F32)
    std = np.maximum(std, F32(eps))
    mask = np.abs((blocks - mean).astype(F32)) > (F32(6.0) * std).astype(F32)
    rows, cols = np.nonzero(mask)
    values = blocks[rows, cols].astype(F32)
    indices = np.stack([rows, cols], axis=1).astype(np.int64)
    blocks = np.where(mask, mean, blocks).astype(F32)
    bmin = blocks.min(axis=1, keepdims=True)
    bmax = blocks.max(axis=1, keepdims=True)
    scales = np.maximum(((bmax - bmin).astype(F32) / levels).astype(F32), F32(eps))
    q = np.clip(np.rint(((blocks - bmin).astype(F32) / scales).astype(F32)), 0, levels).astype(np.uint8)
    per = 8 // bits
    packed = np.zeros((q.shape[0], q.shape[1] // per), dtype=np.uint8)
    for e in range(per):                                     # element 0 of each group in the most significant bits
        packed |= (q[:, e::per] << (bits * (per - 1 - e))).astype(np.uint8)
    return packed, (bmin.astype(F32), scales, values, indices), tuple(x.shape)


def dequantize_bbint(packed: np.ndarray, params, shape, method: str) -> np.ndarray:
    """quantization.py:156-173 / :223-243: code * scale + min, then the outliers are put back."""
    bits = 4 if method == "bbint4" else 2
    per = 8 // bits
    bmin, scales, values, indices = params
    q = np.zeros((packed.shape[0], packed.shape[1] * per), dtype=F32)
    for e in range(per):
        q[:, e::per] = ((packed >> (bits * (per - 1 - e))) & (2 ** bits - 1)).astype(F32)
    out = ((q * scales.astype(F32)).astype(F32) + bmin.astype(F32)).astype(F32)
    if len(values):
        out[indices[:, 0], indices[:, 1]] = values
    return out.reshape(shape)


def dequantize_uniform(codes: np.ndarray, scales: np.ndarray, shape, bits: int) -> np.ndarray:
    """x_hat = (float(code) / levels) * scale, reshaped (quantization.py:105, 295, 306)."""
    vals = (codes.astype(F32) / F32(uniform_levels(bits))) * scales.astype(F32)
    return vals.reshape(shape)


def pack_codes(codes: np.ndarray, bits: int) -> np.ndarray:
    """Offset-binary, MSB-first packing of the row-major code stream into bytes.

    Stored symbol = code + levels (so 2-bit -> {0,1,2}, 4-bit -> 0..14).  Element 0 of every
    group lands in the most significant bits, as in q[0::2]*16 + q[1::2]
    (quantization.py:152) and q[0::4]*64 + ... (quantization.py:217-220).  8-bit codes are
    stored as one offset byte, 16-bit as big-endian offset uint16.  The tail of the last
    byte is zero padded.
    """
    lv = uniform_levels(bits)
    sym = (codes.astype(np.int32).reshape(-1) + lv)
    if bits == 16:
        out = np.empty(sym.size * 2, dtype=np.uint8)
        out[0::2] = (sym >> 8) & 0xFF
        out[1::2] = sym & 0xFF
        return out
    if bits == 8:
        return sym.astype(np.uint8)
    per = 8 // bits
    pad = (-sym.size) % per
    if pad:
        sym = np.concatenate([sym, np.zeros(pad, dtype=np.int32)])
    sym = sym.reshape(-1, per)
    out = np.zeros(sym.shape[0], dtype=np.int32)
    for k in range(per):
        out = (out << bits) | sym[:, k]
    return out.astype(np.uint8)


def unpack_codes(packed: np.ndarray, bits: int, count: int) -> np.ndarray:
    """Inverse of `pack_codes`: returns signed codes (int8 / int16) of length `count`."""
    lv = uniform_levels(bits)
    p = packed.astype(np.int32).reshape(-1)
    if bits == 16:
        sym = (p[0::2] << 8) | p[1::2]
        return (sym[:count] - lv).astype(np.int16)
    if bits == 8:
        return (p[:count] - lv).astype(np.int8)
    per = 8 // bits
    mask = (1 << bits) - 1
    cols = [(p >> (bits * (per - 1 - k))) & mask for k in range(per)]
    sym = np.stack(cols, axis=1).reshape(-1)[:count]
    return (sym - lv).astype(np.int8)


# --------------------------------------------------------------------------- parameters
@dataclass
class OracleParams:
    """Plain mirror of CalderaParams (RCR/caldera/utils/dataclasses.py:11-84)."""
    compute_quantized_component: bool = True
    compute_low_rank_factors: bool = True
    Q_bits: int = 2
    L_bits: int = 2
    R_bits: int = 2
    rank: int = 64
    iters: int = 20
    lplr_iters: int = 5
    activation_aware_LR: bool = True
    update_order: List[str] = field(default_factory=list)
    rand_svd: bool = False
    sigma_reg: float = 0.0


@dataclass
class OracleDecomposition:
    Q: np.ndarray = None
    L: np.ndarray = None
    R: np.ndarray = None
    W: np.ndarray = None
    Q_idxs: np.ndarray = None
    L_idxs: np.ndarray = None
    R_idxs: np.ndarray = None
    Q_scale: object = 1
    L_scale: object = 1
    R_scale: object = 1
    global_scale: float = 1
    errors: Dict[str, List[float]] = field(default_factory=dict)
    best_step: int = -1          # index of the sub-step that produced the returned iterate


# --------------------------------------------------------------------------- Hessian
def hessian_factors(H: np.ndarray, aware: bool, sigma_reg: float):
    """Returns (H_used, H_sqrt, evals, evecs) as the reference builds them (alg.py:44-68).

    Not aware: H_sqrt = H, evals = 1, evecs = H (alg.py:49-51).
    Aware: symmetrise; identity shortcut (alg.py:13); eigh; shift spectrum up to sigma_reg
    when the smallest eigenvalue is below it (alg.py:59-64); H_sqrt = V sqrt(lam) V^T.
    """
    n = H.shape[0]
    H = H.astype(F32)
    if not aware:
        return H, H, np.ones(n, dtype=F32), H
    H = ((H + H.T) / F32(2)).astype(F32)
    eye = np.eye(n, dtype=F32)
    if np.allclose(H, eye, rtol=1e-5, atol=1e-8):
        evals, evecs = np.ones(n, dtype=F32), eye
    else:
        evals, evecs = np.linalg.eigh(H)
        evals, evecs = evals.astype(F32), evecs.astype(F32)
    lo = evals.min()
    if lo < sigma_reg:
        shift = F32(sigma_reg) - lo
        H = (H + shift * eye).astype(F32)
        evals = (evals + shift).astype(F32)
    H_sqrt = (evecs * np.sqrt(evals)[None, :]) @ evecs.T
    return H, H_sqrt.astype(F32), evals, evecs


def weighted_error(W: np.ndarray, H: np.ndarray, Q: np.ndarray, L: np.ndarray, R: np.ndarray) -> float:
    """sqrt( tr(E H E^T) / tr(W H W^T) ), E = Q + L R - W  (alg.py:286-302).

    Written with the same four dense products the reference forms, so that timing this
    function reproduces the reference's cost structure.
    """
    E = (Q + L @ R) - W
    num = np.trace((E @ H) @ E.T)
    den = np.trace((W @ H) @ W.T)
    return float(np.sqrt(num / den))


# --------------------------------------------------------------------------- low rank
def _halko_lowrank(Y: np.ndarray, q: int, niter: int, rng: np.random.Generator):
    """Halko et al. Alg. 5.1 as torch.svd_lowrank runs it (Gaussian test matrix, QR after
    every half step, `niter` power iterations).  Returns U[m,q], S[q], Vh[q,n]."""
    m, n = Y.shape
    omega = rng.standard_normal((n, q)).astype(F32)
    Qm, _ = np.linalg.qr(Y @ omega)
    for _ in range(niter):
        Qn, _ = np.linalg.qr(Y.T @ Qm)
        Qm, _ = np.linalg.qr(Y @ Qn)
    B = Qm.T @ Y
    Ub, S, Vh = np.linalg.svd(B, full_matrices=False)
    return Qm @ Ub, S, Vh


def lowrank_init(residual, H_sqrt, evals, evecs, rank, aware, rand_svd, shape, rng):
    """Closed-form rank-constrained regression (alg.py:201-235)."""
    if aware:
        Y = (residual @ H_sqrt) @ evecs
    else:
        Y = residual
    if rand_svd:
        q = min(rank * 2, min(shape))
        U, S, Vh = _halko_lowrank(Y.astype(F32), q, 2, rng)
    else:
        U, S, Vh = np.linalg.svd(Y.astype(F32), full_matrices=False)
    U, S, Vh = U[:, :rank], S[:rank], Vh[:rank, :]
    if aware:
        L = U
        R = ((S[:, None] * Vh) * (F32(1) / np.sqrt(evals))[None, :]) @ evecs.T
    else:
        rs = np.sqrt(S)
        L = U * rs[None, :]
        R = rs[:, None] * Vh
    return L.astype(F32), R.astype(F32)


def _lstsq(A, B):
    X = np.linalg.lstsq(A, B, rcond=None)[0]
    if np.isnan(X).any():
        X = np.linalg.pinv(A) @ B
    return X.astype(F32)


def _quantize_whole(A: np.ndarray, bits: int):
    """`quantize_matrix` (alg.py:245-250): one scale for the whole tensor."""
    codes, scales, shape = quantize_uniform(A, bits, None)
    return dequantize_uniform(codes, scales, shape, bits), codes, scales


def lplr_refine(residual, H_sqrt, L, R, params: OracleParams):
    """Alternating weighted least squares with quantised factors (alg.py:144-195).

    Returns (L, R, L_codes, R_codes, L_scale, R_scale).  Raises AttributeError when
    lplr_iters == 0, like the reference does on `None.A_idxs` (alg.py:190).
    """
    aware = params.activation_aware_LR
    best = None
    best_err = float("inf")
    RH = residual @ H_sqrt if aware else None
    for _ in range(params.lplr_iters):
        if aware:
            L = _lstsq((R @ H_sqrt).T, RH.T).T
        else:
            L = _lstsq(R.T, residual.T).T
        Lq_t, L_codes, L_scale = _quantize_whole(np.ascontiguousarray(L.T), params.L_bits)
        L = np.ascontiguousarray(Lq_t.T)
        R = _lstsq(L, residual)
        R, R_codes, R_scale = _quantize_whole(R, params.R_bits)
        err = np.linalg.norm((residual - L @ R) @ H_sqrt)
        if err < best_err:
            best_err = err
            best = (L, R, L_codes, R_codes, L_scale, R_scale)
    if best is None:
        raise AttributeError("'NoneType' object has no attribute 'A_idxs'")
    return best


# --------------------------------------------------------------------------- driver
def caldera_oracle(params: OracleParams, W: np.ndarray, H: Optional[np.ndarray] = None,
                   scale_W: bool = True, seed: int = 42) -> OracleDecomposition:
    """Outer alternating minimisation W ~ Q + L R (alg.py:24-112)."""
    W = np.asarray(W, dtype=F32)
    m, n = W.shape
    if scale_W:
        global_scale = float(np.sqrt(np.mean(np.square(W), dtype=F32)))
    else:
        global_scale = 1
    W = (W / F32(global_scale)).astype(F32)
    if H is None:
        H = np.eye(n, dtype=F32)
    H = np.asarray(H, dtype=F32)
    if H.ndim == 1:
        H = np.diag(H)
    H, H_sqrt, evals, evecs = hessian_factors(H, params.activation_aware_LR, params.sigma_reg)
    rng = np.random.default_rng(seed)

    cur = OracleDecomposition(Q=np.zeros((m, n), F32), L=np.zeros((m, params.rank), F32),
                              R=np.zeros((params.rank, n), F32), W=W)
    best = copy.deepcopy(cur)
    errors: Dict[str, List[float]] = {k: [] for k in params.update_order}
    updated = {k: False for k in params.update_order}
    min_error = float("inf")
    step = 0
    for _ in range(params.iters):
        for which in params.update_order:
            if which == "LR" and params.compute_low_rank_factors:
                residual = W - cur.Q
                L, R = lowrank_init(residual, H_sqrt, evals, evecs, params.rank,
                                    params.activation_aware_LR, params.rand_svd, (m, n), rng)
                if params.L_bits < 16 or params.R_bits < 16:
                    (L, R, cur.L_idxs, cur.R_idxs,
                     cur.L_scale, cur.R_scale) = lplr_refine(residual, H_sqrt, L, R, params)
                cur.L, cur.R = L, R
            elif which == "Q" and params.compute_quantized_component:
                residual = W - cur.L @ cur.R if params.compute_low_rank_factors else W
                cur.Q, cur.Q_idxs, cur.Q_scale = _quantize_whole(residual.astype(F32), params.Q_bits)
            updated[which] = True
            err = weighted_error(W, H, cur.Q, cur.L, cur.R)
            errors[which].append(err)
            if err < min_error and all(updated.values()):
                min_error = err
                best = copy.deepcopy(cur)
                best.best_step = step
            step += 1
    best.errors = errors
    best.global_scale = global_scale
    return best
