"""Parity reporting helpers: exact-match fraction of integer codes against a stored reference run.

The north star asks for the code-mismatch rate next to every timing: packed codes must be bit-exact wherever the
quantiser input agrees (iterate 0: the first Q update quantises W / global_scale itself) and the mismatch rate is
reported otherwise (later iterates quantise W - L R, and L R comes from a different -- randomized, bf16 --
factorisation than the reference's exact SVD).  tests/golden/fullsize_*.json / *_codes.npz hold the reference's
results (tests/golden/make_golden_fullsize.py); these helpers compare a CalderaDecomposition with them."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
import torch

from . import _lib

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_fullsize_golden(name: str):
    """(scalars dict, packed best-iterate Q codes as a uint8 numpy array or None) of tests/golden/fullsize_<name>."""
    with open(os.path.join(GOLDEN_DIR, f"fullsize_{name}.json")) as f:
        z = json.load(f)
    path = os.path.join(GOLDEN_DIR, f"fullsize_{name}_codes.npz")
    packed = np.load(path)["q_packed_best"] if os.path.exists(path) else None
    return z, packed


def unpack(packed: torch.Tensor, numel: int, bits: int) -> torch.Tensor:
    """Packed codes (device uint8) -> int8 / int16 codes through the library's own unpack kernel."""
    lib = _lib.load()
    out = torch.empty(numel, dtype=torch.int8 if bits <= 8 else torch.int16, device=packed.device)
    with torch.cuda.device(packed.device):
        _lib.check(lib.cb_unpack_codes(_lib.ptr(packed), numel, bits, _lib.ptr(out), _lib.stream_ptr()), "unpack")
    return out


def code_match_fraction(packed: torch.Tensor, golden_packed: np.ndarray, numel: int, bits: int) -> float:
    """Fraction of the `numel` codes that are identical in the two packed arrays."""
    g = torch.from_numpy(np.ascontiguousarray(golden_packed)).to(packed.device)
    if torch.equal(packed.reshape(-1), g.reshape(-1)):
        return 1.0
    a, b = unpack(packed.reshape(-1), numel, bits), unpack(g.reshape(-1), numel, bits)
    return float((a == b).double().mean().item())


def codes_sha256(codes: torch.Tensor) -> str:
    return hashlib.sha256(codes.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def parity_report(dec, golden_name: str, dec_iter0=None) -> dict:
    """{"code_match_best", "code_match_iter0", "best_err_rel_diff", ...} of a decomposition against the stored
    reference run.  `dec_iter0`: a decomposition of the same layer stopped after the first Q update (iters=1,
    update_order=["Q"]) with the reference's global_scale injected, for the bit-exactness check of iterate 0."""
    z, gp = load_fullsize_golden(golden_name)
    flat = [e for pair in zip(dec.errors["Q"], dec.errors["LR"]) for e in pair]
    best = min(flat[1:])
    rep = {"golden": f"fullsize_{golden_name}", "best_err": best, "best_err_reference": z["best_error"],
           "best_err_rel_diff": abs(best - z["best_error"]) / z["best_error"],
           "q_scale_rel_diff": abs(float(dec.Q_scale.reshape(-1)[0]) - z["Q_scale"]) / z["Q_scale"]}
    if gp is not None and getattr(dec, "Q_packed", None) is not None:
        rep["code_match_best"] = code_match_fraction(dec.Q_packed, gp, int(gp.size) * 4, 2)
    if dec_iter0 is not None:
        rep["code_match_iter0"] = 1.0 if codes_sha256(dec_iter0.Q_idxs.reshape(-1)) == z["q_idxs_iter0_sha256"] else \
            "sha256 mismatch"
        rep["q_scale_iter0_equal"] = bool(np.float32(float(dec_iter0.Q_scale.reshape(-1)[0])) == np.float32(z["Q_scale_iter0"]))
    return rep
