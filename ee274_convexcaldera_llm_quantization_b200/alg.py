"""Drop-in for RCR/caldera/decomposition/alg.py: `caldera()` on the B200 kernel library.

The Python here only validates arguments, allocates outputs/workspace with torch and makes
ONE call into `cb_caldera_layer` (csrc/driver.cu), which enqueues the whole alternating
minimisation on the current CUDA stream.  The single host synchronisation per layer is
the read-back of the error trajectory at the end.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .params import CalderaParams, CalderaDecomposition, QuantInfo  # noqa: F401  (re-exported like the reference)
from .quantization import QuantizerFactory, LowMemoryQuantizer, AbstractQuantizer  # noqa: F401
from .runner import CalderaLayerRunner, workspace_bytes

import os
import threading

# Scratch arenas are reused across calls (one per device and calling thread): a layer needs
# ~0.5 GiB of workspace at 4096 x 4096 and re-allocating it per call costs more than the H2D copy.
_WS_CACHE = {}


_WS_CACHE_MAX_BYTES = int(os.environ.get("CB_WS_CACHE_MAX_BYTES", str(8 << 30)))


def _cached_workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    key = (dev.index, threading.get_ident())
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        _WS_CACHE.pop(key, None)
        # bounded by bytes, not by entries: arenas of threads that are gone are dropped oldest first
        while _WS_CACHE and sum(t.numel() for t in _WS_CACHE.values()) + nbytes > _WS_CACHE_MAX_BYTES:
            _WS_CACHE.pop(next(iter(_WS_CACHE)))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if nbytes <= _WS_CACHE_MAX_BYTES:
            _WS_CACHE[key] = ws
    return ws


def release_workspaces() -> None:
    """Drops the cached scratch arenas and captured graphs (otherwise kept for the life of the process)."""
    from .engine import release_engines
    _WS_CACHE.clear()
    release_engines()


# Captured CUDA graphs of the whole layer live in the per-device LayerEngine (engine.py): slots of
# stream + workspace arena + graphs, LRU-bounded by bytes; caldera(use_cuda_graph=True) and caldera_async()
# submit to it.
_ORDER_CODE = {"Q": 0, "LR": 1}

def _resolve_device(device, W: torch.Tensor) -> torch.device:
    dev = torch.device(device) if device is not None else W.device
    if dev.type != "cuda":
        raise RuntimeError(
            "caldera(): the B200 path runs on CUDA devices only and has no CPU fallback "
            f"(got device={device!r}); use the reference implementation for CPU runs")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _classify_hessian(H: Optional[torch.Tensor], n: int, dev: torch.device):
    """Returns (h_kind, tensor-or-None).  A dense n x n H whose off-diagonal entries are all
    zero (what main.py:163-165 builds with diag_embed) is routed to the diagonal path."""
    if H is None:
        return _lib.CB_H_IDENTITY, None
    H = H.to(dev, torch.float32)
    if H.dim() == 1:
        if H.numel() != n:
            raise ValueError(f"diagonal Hessian has {H.numel()} entries, expected {n}")
        return _lib.CB_H_DIAG, H.contiguous()
    if H.dim() != 2 or H.shape[0] != n or H.shape[1] != n:
        raise ValueError(f"H must be ({n}, {n}) or ({n},), got {tuple(H.shape)}")
    H = H.contiguous()
    lib = _lib.load()
    diag = torch.empty(n, dtype=torch.float32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.cb_hessian_probe(_lib.ptr(H), n, _lib.ptr(diag), _lib.ptr(flag), _lib.stream_ptr()),
                   "hessian_probe")
    if int(flag.item()) == 1:
        return _lib.CB_H_DIAG, diag
    return _lib.CB_H_DENSE, H


def make_c_params(quant_params: CalderaParams, scale_W: bool, global_scale: Optional[float] = None,
                  sketch_width: int = 0, power_iters: int = -1, warm_start: bool = True,
                  seed: int = 0, use_tensor_cores: bool = True, power_iters_warm: int = -1) -> _lib.cb_caldera_params:
    order = list(quant_params.update_order)
    for name in order:
        if name not in _ORDER_CODE:
            raise ValueError(f"update_order entries must be 'Q' or 'LR', got {name!r}")
    if len(order) > 8:
        raise ValueError("update_order longer than 8 entries is not supported")
    p = _lib.cb_caldera_params()
    p.compute_q = int(bool(quant_params.compute_quantized_component))
    p.compute_lr = int(bool(quant_params.compute_low_rank_factors))
    p.q_bits, p.l_bits, p.r_bits = int(quant_params.Q_bits), int(quant_params.L_bits), int(quant_params.R_bits)
    p.rank = int(quant_params.rank)
    p.iters = int(quant_params.iters)
    p.lplr_iters = int(quant_params.lplr_iters)
    p.aware = int(bool(quant_params.activation_aware_LR))
    p.n_order = len(order)
    for i, name in enumerate(order):
        p.order[i] = _ORDER_CODE[name]
    p.rand_svd = int(bool(quant_params.rand_svd))
    p.sigma_reg = float(quant_params.sigma_reg)
    p.scale_w = int(bool(scale_W))
    p.global_scale_in = float(global_scale) if global_scale is not None else 0.0
    p.q_block = 0          # quantize_matrix forces one block per tensor (alg.py:247)
    p.sketch_width = int(sketch_width)
    p.power_iters = int(power_iters)
    p.power_iters_warm = int(power_iters_warm)
    p.warm_start = int(bool(warm_start))
    p.use_tensor_cores = int(bool(use_tensor_cores))
    p.exec_mode = _lib.execution_mode_code()
    p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return p


def _check_factories(quant_params: CalderaParams):
    for fac in (quant_params.quant_factory_Q, quant_params.quant_factory_LR):
        method = str(getattr(fac, "method", "uniform")).lower()
        if method != "uniform":
            raise NotImplementedError(
                f"Quantization method '{method}' not implemented in the B200 path "
                "(the CALDERA hot path is pinned to method='uniform').")


def _validate(quant_params: CalderaParams, W: torch.Tensor):
    if len(W.shape) != 2:
        raise ValueError(f"Support only for 2D matrix, but your input has {len(W.shape)} dimensions.")
    _check_factories(quant_params)
    quant_factors = bool(quant_params.compute_low_rank_factors) and (quant_params.L_bits < 16 or quant_params.R_bits < 16)
    if quant_params.compute_quantized_component:
        assert quant_params.Q_bits in (2, 4, 8, 16), "Bit-width not supported!"
    if quant_factors:
        assert quant_params.L_bits in (2, 4, 8, 16) and quant_params.R_bits in (2, 4, 8, 16), "Bit-width not supported!"
        if quant_params.lplr_iters < 1 and "LR" in quant_params.update_order and quant_params.iters > 0:
            # the reference dereferences best_L_quant_out = None (alg.py:190)
            raise AttributeError("'NoneType' object has no attribute 'A_idxs'")
    return quant_factors


def _build_decomposition(quant_params, p, m, n, scale_W, quant_factors, host, nsteps, nerr_pad, t, W_copy, dev):
    """CalderaDecomposition from the layer's host record (errors | scalars | scales) and its device tensors `t`."""
    errs = host[:nsteps].tolist()
    scal = host[nerr_pad:nerr_pad + 8]
    best_step = int(scal[2].item())
    errors = {name: [] for name in quant_params.update_order}
    k = 0
    for _ in range(p.iters):
        for name in quant_params.update_order:
            errors[name].append(errs[k])
            k += 1
    taken = best_step >= 0
    f32 = dict(dtype=torch.float32, device=dev)
    dec = CalderaDecomposition(Q=t.get("Q"), L=t.get("L"), R=t.get("R"))
    dec.scaleWH = None
    dec.SU = torch.ones(n, **f32)
    dec.SV = torch.ones(m, **f32)
    if taken and p.compute_q:
        dec.Q_idxs = t.get("Q_idxs")
        dec.Q_scale = t["Q_scale"].reshape(1, 1)
        dec.Q_packed = t.get("Q_packed")
    if taken and quant_factors:
        dec.L_idxs, dec.R_idxs = t.get("L_idxs"), t.get("R_idxs")
        dec.L_scale, dec.R_scale = t["L_scale"].reshape(1, 1), t["R_scale"].reshape(1, 1)
        dec.L_packed, dec.R_packed = t.get("L_packed"), t.get("R_packed")
    Wkeep = t.get("W")
    if W_copy == "cpu":
        dec.W = Wkeep.cpu()
    elif W_copy == "device":
        dec.W = Wkeep
    else:
        dec.W = None
    dec.errors = errors
    dec.global_scale = float(scal[0].item()) if scale_W else 1
    dec.best_step = best_step
    stats = scal[4:8].view(torch.int32).tolist()
    dec.device_stats = {"lplr_updates_without_iterate": stats[0], "cholesky_retries": stats[1], "jacobi_sweeps": stats[2],
                        "tc_watchdog": stats[3]}
    if stats[3] != 0:
        raise _lib.CalderaRuntimeError(2001, "caldera: tcgen05 pipeline watchdog fired")
    if stats[1] >= 99:
        # the analogue of the reference's pinv fallback (alg.py:164-165) ran out: the normal equations stayed
        # indefinite after the three ridge retries, so L / R hold garbage
        raise _lib.CalderaRuntimeError(2000, "caldera: Cholesky factorisation failed even after the ridge retries "
                                             "(rank-deficient or non-finite factors)")
    if stats[0] != 0:
        # every inner LPLR error was NaN: the reference dereferences best_L_quant_out = None here (alg.py:190)
        raise AttributeError("'NoneType' object has no attribute 'A_idxs'")
    return dec


_OUT_FIELDS = ("Q", "L", "R", "Q_idxs", "L_idxs", "R_idxs", "Q_packed", "L_packed", "R_packed", "Q_scale", "L_scale", "R_scale")
_DENSE_FIELDS = ("Q", "Q_idxs", "L_idxs", "R_idxs")


def caldera_async(
    quant_params: CalderaParams,
    W: torch.Tensor,
    H: torch.Tensor = None,
    device: str = "cuda",
    use_tqdm: bool = True,
    scale_W: bool = True,
    *,
    W_copy: str = "none",
    global_scale: Optional[float] = None,
    sketch_width: int = 0,
    power_iters: int = -1,
    power_iters_warm: int = -1,
    warm_start: bool = True,
    seed: int = 0,
    return_packed: bool = True,
    use_tensor_cores: bool = True,
    return_dense: bool = True,
    consume=None,
    slots: Optional[int] = None,
    batch: Optional[int] = None,
    batch_hint: Optional[int] = None,
):
    """caldera() without the wait: enqueues the layer on the device's LayerEngine (engine.py) and returns a handle
    whose `.result()` is the CalderaDecomposition.  One host thread can keep dozens of layers in flight this way
    (the engine only blocks the caller when all of its slots are busy).  W / H may be pinned host tensors (staged
    by asynchronous copies) or device tensors.  Arguments are caldera()'s; additionally
    consume(run, kept)  optional hook called with the layer's stream current right after the layer was enqueued: it
             may enqueue device-side copies of the runner's outputs (run.Q_packed, run.L, ...) to wherever they
             are going (a wire-format arena, pinned host buffers); with a hook, and return_dense / return_packed
             False, no per-layer clones are made at all;
    slots / batch  geometry of the device's engine: `slots` graph replays in flight, each advancing up to `batch`
             same-shape layers in lock step (defaults 3 x 16; the batched driver needs the tensor-core path, a
             diagonal / identity Hessian and rank <= 192 -- other configurations run one layer per replay);
    batch_hint  how many layers of this shape are about to be submitted (a short job then uses a smaller batch).
    A handle's layer starts when its batch is full, on `engine.flush()`, or when `.result()` is called."""
    del use_tqdm
    quant_factors = _validate(quant_params, W)
    dev = _resolve_device(device, W)
    _lib.load()
    m, n = int(W.shape[0]), int(W.shape[1])
    from .engine import get_engine
    with torch.cuda.device(dev):
        if H is not None and not H.is_cuda and H.dim() == 1:
            # a host-side diagonal (ideally pinned) is staged by the engine's asynchronous copy
            if H.numel() != n:
                raise ValueError(f"diagonal Hessian has {H.numel()} entries, expected {n}")
            h_kind, Hd = _lib.CB_H_DIAG, (H if H.dtype == torch.float32 else H.float())
        else:
            h_kind, Hd = _classify_hessian(H, n, dev)
        p = make_c_params(quant_params, scale_W, global_scale, sketch_width, power_iters, warm_start, 0,
                          use_tensor_cores, power_iters_warm)
        want_w = W_copy != "none"
        keep = [f for f in _OUT_FIELDS if (return_dense or f not in _DENSE_FIELDS) and (return_packed or "packed" not in f)]

        def consume_all(run, kept):
            for f in keep:
                t = getattr(run, f)
                if t is not None:
                    kept[f] = t.clone()
            if want_w:
                kept["W"] = (run.W_scaled if scale_W else run.W_in).clone()
            if consume is not None:
                consume(run, kept)

        def finish(run, host, kept):
            return _build_decomposition(quant_params, p, m, n, scale_W, quant_factors, host, run.nsteps, run.nerr_pad,
                                        kept, W_copy, dev)

        Wsrc = W if W.dtype == torch.float32 else W.float()
        # a consume hook reads the runner's packed outputs even when no packed clones are returned
        return get_engine(dev, slots, batch).submit(p, Wsrc, h_kind, Hd, seed, finish,
                                                    want_packed=return_packed or consume is not None,
                                                    want_w_scaled=want_w and scale_W, consume=consume_all,
                                                    batch_hint=batch_hint, want_dense=return_dense)


def caldera(
    quant_params: CalderaParams,
    W: torch.Tensor,
    H: torch.Tensor = None,
    device: str = "cuda",
    use_tqdm: bool = True,
    scale_W: bool = True,
    *,
    W_copy: str = "cpu",
    global_scale: Optional[float] = None,
    sketch_width: int = 0,
    power_iters: int = -1,
    power_iters_warm: int = -1,
    warm_start: bool = True,
    seed: int = 0,
    return_packed: bool = True,
    use_tensor_cores: bool = True,
    use_cuda_graph: bool = False,
    return_dense: bool = True,
):
    """Runs CALDERA: decomposes W into Q + L R (alg.py:24-112), all arithmetic on `device`.

    Positional/keyword arguments are the reference's.  `H` may additionally be a 1-D tensor
    holding the diagonal of a diagonal Hessian.  Keyword-only extras (all optional):
    W_copy   where CalderaDecomposition.W lives: "cpu" (reference behaviour, alg.py:81),
             "device" or "none";
    global_scale  inject the reference's global_scale instead of recomputing it;
    sketch_width / power_iters / power_iters_warm / warm_start / seed  knobs of the randomized rank-r step
             (power iterations of the first, random-start step -- default 12 -- and of the steps warm-started
             from the previous outer iteration's basis -- default 3; rand_svd=True: 2 and 2);
    use_tensor_cores  bf16 tcgen05 contractions for aligned shapes (default) or fp32 SIMT everywhere;
    use_cuda_graph  replay a captured CUDA graph of the layer through the device's LayerEngine
             (== caldera_async(...).result());
    return_packed  also return bit-packed codes as Q_packed / L_packed / R_packed;
    return_dense  False: leave Q, Q_idxs (and L_idxs / R_idxs) unset -- for callers that only consume the packed
             codes, scales and factors (the model-level job), which saves the m x n fp32 + int8 copies per layer.
    `use_tqdm` is accepted and ignored (the loop runs on the device).
    """
    if use_cuda_graph:
        return caldera_async(quant_params, W, H, device, use_tqdm, scale_W, W_copy=W_copy, global_scale=global_scale,
                             sketch_width=sketch_width, power_iters=power_iters, power_iters_warm=power_iters_warm,
                             warm_start=warm_start, seed=seed, return_packed=return_packed,
                             use_tensor_cores=use_tensor_cores, return_dense=return_dense, batch_hint=1).result()
    quant_factors = _validate(quant_params, W)
    dev = _resolve_device(device, W)
    _lib.load()
    m, n = int(W.shape[0]), int(W.shape[1])
    with torch.cuda.device(dev):
        Wd = W.to(dev, torch.float32, non_blocking=True).contiguous()
        h_kind, Hd = _classify_hessian(H, n, dev)
        p = make_c_params(quant_params, scale_W, global_scale, sketch_width, power_iters, warm_start, seed,
                          use_tensor_cores, power_iters_warm)
        ws = _cached_workspace(workspace_bytes(p, m, n, h_kind), dev)
        run = CalderaLayerRunner(p, m, n, h_kind, dev, want_packed=return_packed,
                                 want_w_scaled=(W_copy != "none"), workspace=ws)
        run.enqueue(Wd, Hd)
        host = run.read_small()                       # the one synchronisation of the layer
        run.ws = None                                 # the arena stays in the per-thread cache
        del ws
        t = {f: getattr(run, f) for f in _OUT_FIELDS
             if getattr(run, f) is not None and (return_dense or f not in _DENSE_FIELDS)}
        for f in ("Q_scale", "L_scale", "R_scale"):   # views into the runner's result record
            if f in t:
                t[f] = t[f].clone()
        t["W"] = run.W_scaled if scale_W else Wd
        return _build_decomposition(quant_params, p, m, n, scale_W, quant_factors, host, run.nsteps, run.nerr_pad,
                                    t, W_copy, dev)


def activation_aware_error(W: torch.Tensor, H: torch.Tensor, caldera_info: CalderaDecomposition, device: str):
    """alg.py:286-302 for a decomposition held as dense tensors: sqrt(tr(E H E^T) / tr(W H W^T))."""
    dev = _resolve_device(device, W)
    lib = _lib.load()
    with torch.cuda.device(dev):
        Wd = W.to(dev, torch.float32).contiguous()
        m, n = Wd.shape
        h_kind, Hd = _classify_hessian(H, n, dev)
        What = (caldera_info.Q + caldera_info.L @ caldera_info.R) * caldera_info.global_scale
        E = (What - Wd).contiguous()
        acc = torch.zeros(2, dtype=torch.float64, device=dev)
        _lib.check(lib.cb_weighted_error(_lib.ptr(E), m, n, None, 8, None, None, None, 0, _lib.ptr(Hd), h_kind,
                                         _lib.ptr(acc[0:1]), None, None, 0, _lib.stream_ptr()), "weighted_error")
        _lib.check(lib.cb_weighted_error(_lib.ptr(Wd), m, n, None, 8, None, None, None, 0, _lib.ptr(Hd), h_kind,
                                         _lib.ptr(acc[1:2]), None, None, 0, _lib.stream_ptr()), "weighted_error")
        num, den = acc.tolist()
    return float((num / den) ** 0.5)
