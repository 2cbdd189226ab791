"""Forward pass straight from the packed decomposition.

The reference reconstructs `W_hat = Q + L @ R` as a dense fp32 matrix and multiplies with that
(main.py:197, README.md:182).  `packed_linear` consumes what `caldera()` / the scheduler's blobs hold --
bit-packed Q codes + one scale, fp32 (or fp16) factors -- without ever materialising Q in HBM."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def packed_linear(x: torch.Tensor, Q_packed: torch.Tensor, Q_scale: torch.Tensor, q_bits: int, out_features: int,
                  L: Optional[torch.Tensor] = None, R: Optional[torch.Tensor] = None, global_scale: float = 1.0,
                  watchdog=True) -> torch.Tensor:
    """y = x @ ((Q + L @ R) * global_scale).T for x of shape (..., in_features); all tensors on one CUDA device.
    Q_packed: uint8, out_features * in_features codes of q_bits each (row-major, MSB-first); Q_scale: 1 element.

    watchdog: the tcgen05 pipeline sets a device flag instead of hanging when one of its bounded barrier waits
    expires, and y is then undefined.  True (default): read the flag back after the call (one host
    synchronisation) and raise CalderaRuntimeError if it is set.  A 1-element int32 CUDA tensor: the kernel ORs
    into it and the caller checks it once per forward pass (`check_watchdog`).  False: not checked."""
    if not x.is_cuda:
        raise ValueError("packed_linear runs on a CUDA device only (there is no CPU fallback)")
    lib = _lib.load()
    n = int(x.shape[-1])
    m = int(out_features)
    x2 = x.reshape(-1, n).to(torch.float32).contiguous()
    T = int(x2.shape[0])
    r = 0 if L is None else int(L.shape[1])
    if Q_packed.numel() != lib.cb_packed_bytes(m * n, int(q_bits)):
        raise ValueError(f"Q_packed holds {Q_packed.numel()} bytes, expected {lib.cb_packed_bytes(m * n, int(q_bits))}")
    with torch.cuda.device(x.device):
        Lc = None if L is None else L.to(x.device, torch.float32).contiguous()
        Rc = None if R is None else R.to(x.device, torch.float32).contiguous()
        qs = Q_scale.to(x.device, torch.float32).reshape(-1)[:1].contiguous()
        y = torch.empty((T, m), dtype=torch.float32, device=x.device)
        own_flag = not torch.is_tensor(watchdog)
        flag = torch.zeros(1, dtype=torch.int32, device=x.device) if own_flag else watchdog
        if not own_flag and (flag.dtype != torch.int32 or flag.device != x.device or flag.numel() < 1):
            raise ValueError("watchdog tensor must be a 1-element int32 tensor on x's device")
        ws = torch.empty(lib.cb_packed_linear_workspace_bytes(T, m, n, r), dtype=torch.uint8, device=x.device)
        st = lib.cb_packed_linear_f32(_lib.ptr(x2), T, n, _lib.ptr(Q_packed.contiguous()), int(q_bits), _lib.ptr(qs),
                                      _lib.ptr(Lc) if Lc is not None else None, _lib.ptr(Rc) if Rc is not None else None,
                                      m, r, float(global_scale), _lib.ptr(y), _lib.ptr(flag), _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr())
        _lib.check(st, "packed_linear")
        if own_flag and watchdog:
            check_watchdog(flag)
    return y.reshape(*x.shape[:-1], m)


def check_watchdog(flag: torch.Tensor) -> None:
    """Raises if the tcgen05 pipeline watchdog fired in any packed_linear call that used `flag` (synchronises)."""
    if int(flag.reshape(-1)[0].item()) != 0:
        raise _lib.CalderaRuntimeError(2001, "packed_linear: tcgen05 pipeline watchdog fired; the output is undefined")
