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
_WARMED: set = set()             # configurations that already ran eagerly once on a device (BatchRunner.capture)
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


# ----------------------------------------------------------------------------- batched layers
class _LayerView:
    """One layer of a BatchRunner, with the attribute names of CalderaLayerRunner (views into the layer's slab)."""
    __slots__ = ("Q", "L", "R", "Q_idxs", "L_idxs", "R_idxs", "Q_packed", "L_packed", "R_packed", "Q_scale", "L_scale",
                 "R_scale", "W_scaled", "W_in", "h_in", "small", "nsteps", "nerr_pad", "graph_kernels")


class BatchRunner:
    """`batch` same-shape layers decomposed in lock step by ONE captured CUDA graph of cb_caldera_batch.

    Every per-layer buffer (input staging, outputs, result record, seed, workspace) lives in that layer's slab;
    slabs are `stride` bytes apart inside one allocation, which is what lets a single kernel launch serve the whole
    batch (csrc/driver.cu).  `slab` may be passed in (and shared with the runners of other shapes that are never
    in flight at the same time)."""

    def __init__(self, c_params: "_lib.cb_caldera_params", m: int, n: int, h_kind: int, batch: int, device: torch.device,
                 want_packed: bool = True, want_w_scaled: bool = False, slab: Optional[torch.Tensor] = None,
                 want_dense: bool = True):
        self.lib = _lib.load()
        self.p = c_params
        p = c_params
        self.m, self.n, self.h_kind, self.device, self.batch = int(m), int(n), int(h_kind), device, int(batch)
        r = int(p.rank)
        self.quant_factors = bool(p.compute_lr) and (p.l_bits < 16 or p.r_bits < 16)
        self.nsteps = p.iters * p.n_order
        self.nerr_pad = (self.nsteps + 3) // 4 * 4
        self.nsmall = self.nerr_pad + 8 + 12
        self.off, self.stride, self.ws_bytes = self._layout(p, m, n, h_kind, want_packed, want_w_scaled, want_dense)
        total = self.stride * self.batch
        with torch.cuda.device(device):
            if slab is not None and slab.numel() >= total and slab.device == device:
                self.slab = slab
            else:
                self.slab = torch.empty(total, dtype=torch.uint8, device=device)
        base = self.slab.data_ptr()
        assert base % 256 == 0
        self.out = _lib.cb_caldera_out()
        for name in ("Q", "L", "R", "Q_idxs", "Q_packed", "L_idxs", "R_idxs", "L_packed", "R_packed", "W_scaled"):
            setattr(self.out, name, base + self.off[name][0] if name in self.off else None)
        so = self.off["small"][0]
        self.out.errors = base + so
        self.out.scalars = base + so + 4 * self.nerr_pad
        self.out.Q_scale = base + so + 4 * (self.nerr_pad + 8)
        self.out.L_scale = base + so + 4 * (self.nerr_pad + 12)
        self.out.R_scale = base + so + 4 * (self.nerr_pad + 16)
        self.out.seed_dev = base + self.off["seed"][0]
        # strided views over all layers: seeds (int64) and result records (float32)
        self.seeds = torch.as_strided(self.slab.view(torch.int64), (self.batch,), (self.stride // 8,), self.off["seed"][0] // 8)
        self.small_all = torch.as_strided(self.slab.view(torch.float32), (self.batch, self.nsmall), (self.stride // 4, 1), so // 4)
        cd = lambda b: torch.int8 if b <= 8 else torch.int16  # noqa: E731
        shapes = {"W_in": (torch.float32, (m, n)), "h_in": (torch.float32, (n,)), "Q": (torch.float32, (m, n)),
                  "L": (torch.float32, (m, r)), "R": (torch.float32, (r, n)), "Q_idxs": (cd(p.q_bits), (1, m * n)),
                  "L_idxs": (cd(p.l_bits), (1, r * m)), "R_idxs": (cd(p.r_bits), (1, r * n)), "Q_packed": (torch.uint8, None),
                  "L_packed": (torch.uint8, None), "R_packed": (torch.uint8, None), "W_scaled": (torch.float32, (m, n))}
        self.layers = []
        for b in range(self.batch):
            v = _LayerView()
            for name, (dt, shp) in shapes.items():
                if name in self.off:
                    o0, nb = self.off[name]
                    t = self.slab[b * self.stride + o0: b * self.stride + o0 + nb].view(dt)
                    setattr(v, name, t.reshape(shp) if shp is not None else t)
                else:
                    setattr(v, name, None)
            v.small = self.small_all[b]
            k = self.nerr_pad + 8
            v.Q_scale, v.L_scale, v.R_scale = v.small[k:k + 1], v.small[k + 4:k + 5], v.small[k + 8:k + 9]
            v.nsteps, v.nerr_pad = self.nsteps, self.nerr_pad
            self.layers.append(v)
        self.graph = None
        self.graph_kernels = 0
        self.seeds_host = torch.zeros(self.batch, dtype=torch.int64).pin_memory()

    @staticmethod
    def _layout(p, m, n, h_kind, want_packed, want_w_scaled, want_dense=True):
        """Byte offsets of one layer's buffers inside its slab: ({name: (offset, nbytes)}, stride, workspace bytes)."""
        lib = _lib.load()
        r = int(p.rank)
        quant_factors = bool(p.compute_lr) and (p.l_bits < 16 or p.r_bits < 16)
        nsmall = (p.iters * p.n_order + 3) // 4 * 4 + 8 + 12
        cb = lambda b: 1 if b <= 8 else 2  # noqa: E731
        ws_bytes = int(lib.cb_caldera_layer_workspace_bytes(C.byref(p), m, n, h_kind)) or 256
        fields = [("W_in", 4 * m * n), ("h_in", 4 * n if h_kind == _lib.CB_H_DIAG else 0), ("small", 4 * nsmall),
                  ("seed", 8), ("Q", 4 * m * n if want_dense else 0), ("L", 4 * m * r), ("R", 4 * r * n), ("Q_idxs", m * n * cb(p.q_bits)),
                  ("Q_packed", int(lib.cb_packed_bytes(m * n, p.q_bits)) if want_packed else 0),
                  ("L_idxs", r * m * cb(p.l_bits) if quant_factors else 0),
                  ("R_idxs", r * n * cb(p.r_bits) if quant_factors else 0),
                  ("L_packed", int(lib.cb_packed_bytes(m * r, p.l_bits)) if (want_packed and quant_factors) else 0),
                  ("R_packed", int(lib.cb_packed_bytes(r * n, p.r_bits)) if (want_packed and quant_factors) else 0),
                  ("W_scaled", 4 * m * n if (p.scale_w and want_w_scaled) else 0), ("ws", ws_bytes)]
        off, o = {}, 0
        for name, nbytes in fields:
            if nbytes > 0:
                off[name] = (o, nbytes)
                o += (nbytes + 255) // 256 * 256
        return off, o, ws_bytes

    @staticmethod
    def slab_stride(p, m, n, h_kind, want_packed=True, want_w_scaled=False, want_dense=True) -> int:
        return BatchRunner._layout(p, m, n, h_kind, want_packed, want_w_scaled, want_dense)[1]

    def _ptr(self, name):
        return C.c_void_p(self.slab.data_ptr() + self.off[name][0]) if name in self.off else None

    def enqueue(self) -> None:
        """Asynchronous: the whole batch on the current stream, reading the staged inputs of every slab."""
        with torch.cuda.device(self.device):
            st = self.lib.cb_caldera_batch(C.byref(self.p), self.batch, self.stride, self._ptr("W_in"), self.m, self.n,
                                           self._ptr("h_in"), self.h_kind, C.byref(self.out), self._ptr("ws"), self.ws_bytes,
                                           _lib.stream_ptr())
        _lib.check(st, "caldera_batch")

    def capture(self) -> None:
        if self.graph is not None:
            return
        with _CAPTURE_LOCK, torch.cuda.device(self.device):
            for v in self.layers:
                v.W_in.normal_(0.0, 0.02)
                if v.h_in is not None:
                    v.h_in.fill_(1.0)
            self.seeds.zero_()
            # one eager pass per (device, configuration, shape): every kernel variant of this configuration is loaded
            # and has its attributes set outside a capture.  Further runners of the same configuration (other slots,
            # other batch sizes) launch exactly those kernels and are captured directly.
            sig = (self.device.index, bytes(self.p), self.m, self.n, self.h_kind, self.batch > 1,
                   "Q_packed" in self.off, "W_scaled" in self.off, "Q" in self.off)
            if sig not in _WARMED:
                self.enqueue()
                torch.cuda.current_stream().synchronize()
                _WARMED.add(sig)
            g = torch.cuda.CUDAGraph()
            n0 = self.lib.cb_kernel_launch_count()
            with torch.cuda.graph(g, stream=_capture_stream(self.device), capture_error_mode="thread_local"):
                self.enqueue()
            self.graph_kernels = int(self.lib.cb_kernel_launch_count() - n0)
            for v in self.layers:
                v.graph_kernels = self.graph_kernels / self.batch
            self.graph = g

    def stage(self, b: int, W: torch.Tensor, h: Optional[torch.Tensor]) -> None:
        """Asynchronous copy of one layer's inputs (device or pinned host tensors) into slab b."""
        v = self.layers[b]
        v.W_in.copy_(W, non_blocking=True)
        if v.h_in is not None and h is not None:
            v.h_in.copy_(h, non_blocking=True)

    def replay(self, seeds) -> None:
        """Asynchronous on the current stream: set the per-layer seeds and replay the batch's graph (capture() first:
        capturing overwrites the staged inputs with its warm-up data)."""
        if self.graph is None:
            raise RuntimeError("BatchRunner.replay(): call capture() before staging inputs")
        with torch.cuda.device(self.device):
            # (pinned staging: a pageable source would make the copy wait for everything queued on this stream)
            self.seeds_host[:len(seeds)] = torch.tensor(list(seeds), dtype=torch.int64)
            self.seeds.copy_(self.seeds_host, non_blocking=True)
            self.graph.replay()
            self.lib.cb_note_launches(self.graph_kernels)
