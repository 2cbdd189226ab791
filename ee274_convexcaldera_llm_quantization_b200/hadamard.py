"""Hadamard (incoherence) pre-rotation around caldera().

Mirrors `hadamard_transform(W, inverse=False, original_shape=None)` of the reference driver (main.py:92-133),
which pads W with zeros to the next powers of two, builds dense normalised Hadamard matrices with scipy and
forms H1 @ W_padded @ H2 on the CPU in fp64.  Here the same transform is a fast Walsh-Hadamard transform on
the GPU in fp32."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def next_power_of_two(n: int) -> int:
    """main.py:75-77."""
    return 1 << (int(n) - 1).bit_length()


def hadamard_transform(W: torch.Tensor, inverse: bool = False, original_shape: Optional[Tuple[int, int]] = None):
    """Forward: returns (H1 @ pad(W) @ H2, (rows, cols)).  Inverse: W is the padded-size matrix; returns the
    leading `original_shape` block of H1 @ W @ H2 (H is symmetric and orthogonal, main.py:124-129)."""
    if inverse and original_shape is None:
        raise ValueError("original_shape is required for the inverse transform")
    if not W.is_cuda:
        raise ValueError("hadamard_transform runs on a CUDA device only (there is no CPU fallback)")
    lib = _lib.load()
    rows, cols = (int(W.shape[0]), int(W.shape[1])) if not inverse else (int(original_shape[0]), int(original_shape[1]))
    prows, pcols = next_power_of_two(rows), next_power_of_two(cols)
    Wc = W.to(torch.float32).contiguous()
    if inverse and tuple(Wc.shape) != (prows, pcols):
        raise ValueError(f"inverse transform expects the padded {prows} x {pcols} matrix, got {tuple(Wc.shape)}")
    with torch.cuda.device(W.device):
        out = torch.empty((prows, pcols), dtype=torch.float32, device=W.device)
        ws = torch.empty(lib.cb_hadamard_workspace_bytes(prows, pcols), dtype=torch.uint8, device=W.device)
        st = lib.cb_hadamard_transform_f32(_lib.ptr(Wc), int(Wc.shape[0]), int(Wc.shape[1]), _lib.ptr(out), prows, pcols,
                                           _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
        _lib.check(st, "hadamard_transform")
    if inverse:
        return out[:rows, :cols].contiguous()
    return out, (rows, cols)
