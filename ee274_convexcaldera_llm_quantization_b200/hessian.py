"""Hessian accumulation from calibration activations.

The reference accumulates `activations @ activations.T` in fp64 on the CPU, one calibration sample at a
time (main.py:302-319), and passes `diag_embed(h)` or the dense matrix on to `caldera()`.  This keeps the
running sums on the GPU: `HessianAccumulator.add(X)` with X of shape (tokens, in_features)."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


class HessianAccumulator:
    """Running H = sum_batches X^T X (dense, optional) and h = diag(H) on `device`; `finalize()` divides by the
    number of batches (the reference's per-sample mean, main.py:314) or by `count` if given."""

    def __init__(self, in_features: int, device="cuda", dense: bool = True):
        self.n = int(in_features)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("HessianAccumulator runs on a CUDA device only (there is no CPU fallback)")
        self.H = torch.zeros(self.n, self.n, dtype=torch.float32, device=self.device) if dense else None
        self.h = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.batches = 0
        self._ws = None
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    def add(self, X: torch.Tensor) -> None:
        lib = _lib.load()
        X2 = X.reshape(-1, X.shape[-1]).to(self.device, torch.float32).contiguous()
        if X2.shape[1] != self.n:
            raise ValueError(f"activations have {X2.shape[1]} features, expected {self.n}")
        T = int(X2.shape[0])
        with torch.cuda.device(self.device):
            need = lib.cb_hessian_accumulate_workspace_bytes(T, self.n) if self.H is not None else 0
            if need and (self._ws is None or self._ws.numel() < need):
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            st = lib.cb_hessian_accumulate_f32(_lib.ptr(X2), T, self.n, _lib.ptr(self.H) if self.H is not None else None,
                                               _lib.ptr(self.h), _lib.ptr(self._flag),
                                               _lib.ptr(self._ws) if need else None, need, _lib.stream_ptr())
            _lib.check(st, "hessian_accumulate")
        self.batches += 1

    def finalize(self, count: Optional[int] = None):
        """Returns (H or None, h) divided by `count` (default: number of batches added)."""
        c = float(count if count is not None else max(self.batches, 1))
        if int(self._flag.item()) != 0:
            raise _lib.CalderaRuntimeError(2001, "hessian_accumulate: tcgen05 pipeline watchdog fired")
        return (self.H / c if self.H is not None else None), self.h / c
