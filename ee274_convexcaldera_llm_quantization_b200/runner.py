"""Pre-planned, re-usable launcher for cb_caldera_layer.

`caldera()` builds one of these per call; the layer-sharded scheduler and bench.py keep one
per distinct layer shape so that outputs and the (hundreds of MiB) workspace are allocated
once and every layer is a single asynchronous enqueue on the current stream."""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional

import torch

from . import _lib


def workspace_bytes(c_params, m: int, n: int, h_kind: int) -> int:
    return int(_lib.load().cb_caldera_layer_workspace_bytes(C.byref(c_params), m, n, h_kind)) or 256


# stream capture is a process-wide affair for torch's allocator and RNG bookkeeping: one at a time
_CAPTURE_LOCK = threading.Lock()
# Captures run on a dedicated high-priority stream per device.  torch hands out ordinary streams from a
# pool of 32 per device and priority, round robin, so the default capture stream of torch.cuda.graph can
# be the very stream another worker thread is using (and synchronising) -- which invalidates the capture.
_CAPTURE_STREAMS = {}


def _capture_stream(device: torch.device) -> "torch.cuda.Stream":
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _CAPTURE_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device, priority=-1)
        _CAPTURE_STREAMS[key] = st
    return st


class CalderaLayerRunner:
    def __init__(self, c_params: "_lib.cb_caldera_params", m: int, n: int, h_kind: int, device: torch.device,
                 want_packed: bool = True, want_w_scaled: bool = True, workspace: Optional[torch.Tensor] = None):
        self.lib = _lib.load()
        self.p = c_params
        self.m, self.n, self.h_kind, self.device = int(m), int(n), int(h_kind), device
        p = c_params
        r = int(p.rank)
        self.quant_factors = bool(p.compute_lr) and (p.l_bits < 16 or p.r_bits < 16)
        f32 = dict(dtype=torch.float32, device=device)
        self.nsteps = p.iters * p.n_order
        with torch.cuda.device(device):
            self.Q = torch.empty((m, n), **f32)
            self.L = torch.empty((m, r), **f32)
            self.R = torch.empty((r, n), **f32)
            # errors | 8 scalars | Q/L/R scales (padded so each section starts 16-byte aligned)
            self.nerr_pad = (self.nsteps + 3) // 4 * 4
            self.small = torch.zeros(self.nerr_pad + 8 + 12, **f32)
            self.errors_d = self.small[:self.nsteps]
            self.scalars_d = self.small[self.nerr_pad:self.nerr_pad + 8]
            o = self.nerr_pad + 8
            self.Q_scale, self.L_scale, self.R_scale = (self.small[o + 4 * i:o + 4 * i + 1] for i in range(3))
            cd = lambda b: torch.int8 if b <= 8 else torch.int16  # noqa: E731
            self.Q_idxs = torch.empty((1, m * n), dtype=cd(p.q_bits), device=device) if p.compute_q else None
            self.L_idxs = torch.empty((1, r * m), dtype=cd(p.l_bits), device=device) if self.quant_factors else None
            self.R_idxs = torch.empty((1, r * n), dtype=cd(p.r_bits), device=device) if self.quant_factors else None
            self.Q_packed = self.L_packed = self.R_packed = None
            if want_packed and p.compute_q:
                self.Q_packed = torch.empty(self.lib.cb_packed_bytes(m * n, p.q_bits), dtype=torch.uint8, device=device)
            if want_packed and self.quant_factors:
                self.L_packed = torch.empty(self.lib.cb_packed_bytes(m * r, p.l_bits), dtype=torch.uint8, device=device)
                self.R_packed = torch.empty(self.lib.cb_packed_bytes(r * n, p.r_bits), dtype=torch.uint8, device=device)
            self.W_scaled = torch.empty((m, n), **f32) if (p.scale_w and want_w_scaled) else None
            self.ws_bytes = int(self.lib.cb_caldera_layer_workspace_bytes(C.byref(p), m, n, h_kind)) or 256
            if workspace is not None and workspace.numel() >= self.ws_bytes and workspace.device == device:
                self.ws = workspace
            else:
                self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=device)
        self.out = _lib.cb_caldera_out()
        for name in ("Q", "L", "R", "Q_idxs", "Q_scale", "Q_packed", "L_idxs", "R_idxs", "L_scale", "R_scale",
                     "L_packed", "R_packed", "W_scaled"):
            t = getattr(self, name)
            setattr(self.out, name, None if t is None else t.data_ptr())
        # raw addresses: data_ptr() of an empty slice (update_order == []) is null
        self.out.errors = self.small.data_ptr()
        self.out.scalars = self.small.data_ptr() + 4 * self.nerr_pad
        # per-layer seed lives in device memory so that a captured graph can be replayed with any seed
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=device)
        self.out.seed_dev = self.seed_dev.data_ptr()
        self.graph = None
        self.W_in = None
        self.h_in = None

    def enqueue(self, W: torch.Tensor, h: Optional[torch.Tensor]) -> None:
        """Asynchronous: enqueues the whole layer on the current stream of `device`."""
        assert W.is_cuda and W.dtype == torch.float32 and W.is_contiguous() and tuple(W.shape) == (self.m, self.n)
        with torch.cuda.device(self.device):
            st = self.lib.cb_caldera_layer(C.byref(self.p), _lib.ptr(W), self.m, self.n, _lib.ptr(h), self.h_kind,
                                           C.byref(self.out), _lib.ptr(self.ws), self.ws_bytes, _lib.stream_ptr())
        _lib.check(st, "caldera")

    # ---- CUDA-graph replay: one graph launch per layer instead of ~600 kernel launches
    def capture(self) -> None:
        """Captures the whole layer into a CUDA graph reading from runner-owned input buffers."""
        if self.graph is not None:
            return
        with _CAPTURE_LOCK, torch.cuda.device(self.device):
            self.W_in = torch.empty((self.m, self.n), dtype=torch.float32, device=self.device)
            self.h_in = None
            if self.h_kind == _lib.CB_H_DIAG:
                self.h_in = torch.ones(self.n, dtype=torch.float32, device=self.device)
            elif self.h_kind == _lib.CB_H_DENSE:
                self.h_in = torch.eye(self.n, dtype=torch.float32, device=self.device)
            self.W_in.normal_(0.0, 0.02)
            self.enqueue(self.W_in, self.h_in)          # eager warm-up: one-time attribute setup happens outside capture
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = self.lib.cb_kernel_launch_count()
            with torch.cuda.graph(g, stream=_capture_stream(self.device), capture_error_mode="thread_local"):
                self.enqueue(self.W_in, self.h_in)
            self.graph_kernels = int(self.lib.cb_kernel_launch_count() - n0)   # kernel nodes in the graph
            self.graph = g

    def launch(self, W: torch.Tensor, h: Optional[torch.Tensor], seed: int = 0) -> None:
        """Asynchronous on the current stream: stage the inputs (device or pinned host tensors) into the
        graph's input buffers and replay it."""
        if self.graph is None:
            self.capture()
        with torch.cuda.device(self.device):
            self.W_in.copy_(W, non_blocking=True)
            if self.h_in is not None and h is not None:
                self.h_in.copy_(h, non_blocking=True)
            self.seed_dev.fill_(int(seed))
            self.graph.replay()
            self.lib.cb_note_launches(self.graph_kernels)

    def read_small(self) -> torch.Tensor:
        """The one host synchronisation of a layer: error trajectory, scalars, scales."""
        return self.small.cpu()
