"""Asynchronous layer engine: many decompositions in flight on one GPU, driven by ONE host thread.

The reference decomposes layers one after the other and blocks on every one (main.py:147-199; host syncs
at alg.py:39, :59, :300).  A layer here is a single CUDA-graph replay (runner.py) whose only latency-bound part
is a chain of single-CTA factorisation kernels, so throughput comes from keeping many independent layers in
flight.  Round 1 did that with one blocked host thread per layer in flight, which made the end-to-end rate
`threads / layer latency` and capped it by the host's core count.  The engine decouples the two:

  submit()   picks a free slot (stream + workspace arena + captured graph for this shape/parameters), enqueues
             H2D staging of W and h, the graph replay, the caller's device-side `consume` copies (e.g. straight
             into a wire-format arena or pinned host buffers) and the D2H copy of the ~100-byte result record,
             records an event and returns a LayerHandle at once.  It only blocks when every slot is busy.
  LayerHandle.result()  waits for that event and builds the CalderaDecomposition.

Slots own their stream and workspace; the graphs of different shapes captured in one slot share its arena, so
device memory is slots x (largest layer footprint), whatever the number of distinct shapes.  Captured runners are
kept per slot in LRU order under a byte budget (`CB_ENGINE_MAX_BYTES`, default 70 % of the device memory that is
free when the engine is created).
"""
from __future__ import annotations

import os
import threading
from collections import OrderedDict
from typing import Callable, Optional

import torch

from . import _lib
from .runner import CalderaLayerRunner, workspace_bytes


def _params_signature(p) -> tuple:
    return tuple(getattr(p, name) if name != "order" else tuple(p.order)
                 for name, _ in p._fields_ if name != "seed")


def _tensor_bytes(*tensors) -> int:
    return sum(t.numel() * t.element_size() for t in tensors if t is not None)


class LayerHandle:
    """One layer in flight.  `result()` blocks until it is done and returns what `finish` builds."""

    def __init__(self, engine: "LayerEngine", slot: "_Slot", run: CalderaLayerRunner, finish: Callable):
        self._engine, self._slot, self.run, self._finish = engine, slot, run, finish
        self._host = None
        self._value = None
        self._done = False
        self.kept = {}           # device tensors produced by `consume` that belong to this layer

    def _collect(self) -> None:
        """Waits for the layer and frees its slot (idempotent; called by result() or when the slot is reused)."""
        if self._host is not None:
            return
        slot = self._slot
        slot.event.synchronize()
        self._host = slot.host_small[:self.run.small.numel()].clone()
        slot.handle = None
        self._engine._release(slot)

    def done(self) -> bool:
        return self._host is not None or self._slot.event.query()

    def result(self):
        if not self._done:
            self._collect()
            self._value = self._finish(self.run, self._host, self.kept)
            self._done = True
            self._finish = None
        return self._value


class _Slot:
    def __init__(self, device: torch.device, index: int):
        self.index = index
        self.stream = torch.cuda.Stream(device=device)
        self.event = torch.cuda.Event()
        self.ws: Optional[torch.Tensor] = None
        self.runners: "OrderedDict[tuple, CalderaLayerRunner]" = OrderedDict()
        self.host_small = torch.empty(4096, dtype=torch.float32).pin_memory()
        self.handle: Optional[LayerHandle] = None


class LayerEngine:
    def __init__(self, device: torch.device, slots: int = 32, max_bytes: Optional[int] = None):
        self.device = device
        self.lock = threading.RLock()
        with torch.cuda.device(device):
            self.slots = [_Slot(device, i) for i in range(max(1, int(slots)))]
            if max_bytes is None:
                env = os.environ.get("CB_ENGINE_MAX_BYTES")
                max_bytes = int(env) if env else int(0.7 * torch.cuda.mem_get_info(device)[0])
        self.max_bytes = int(max_bytes)
        self.free = list(reversed(self.slots))       # pop() hands out slot 0 first
        self.inflight = []                           # slots in submission order
        self.bytes = 0
        self.kernels_replayed = 0

    # ------------------------------------------------------------------ memory accounting
    @staticmethod
    def _runner_bytes(run: CalderaLayerRunner) -> int:
        return _tensor_bytes(run.Q, run.L, run.R, run.Q_idxs, run.L_idxs, run.R_idxs, run.Q_packed, run.L_packed,
                             run.R_packed, run.W_scaled, run.W_in, run.h_in, run.small)

    def _evict(self, need: int, keep_slot: _Slot) -> None:
        """Drops least-recently-used runners of idle slots until `need` more bytes fit the budget."""
        for slot in self.slots:
            if self.bytes + need <= self.max_bytes:
                return
            if slot.handle is not None and slot is not keep_slot:
                continue
            while slot.runners and self.bytes + need > self.max_bytes:
                _, old = slot.runners.popitem(last=False)
                self.bytes -= self._runner_bytes(old)
                if old.ws is not slot.ws:
                    self.bytes -= _tensor_bytes(old.ws)

    def reserve_workspace(self, nbytes: int) -> None:
        """Sizes every slot's arena for the largest layer of a job up front (graphs captured later share it)."""
        with self.lock, torch.cuda.device(self.device):
            for slot in self.slots:
                if slot.ws is None or slot.ws.numel() < nbytes:
                    if slot.handle is not None:
                        slot.handle._collect()
                    self._grow(slot, nbytes)

    def _grow(self, slot: _Slot, nbytes: int) -> None:
        # runners captured against the old arena keep it alive through their own reference and stay valid
        if slot.ws is not None and not any(r.ws is slot.ws for r in slot.runners.values()):
            self.bytes -= _tensor_bytes(slot.ws)
        self._evict(nbytes, slot)
        slot.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.bytes += nbytes

    # ------------------------------------------------------------------ slots
    def _acquire(self) -> _Slot:
        if not self.free:
            oldest = self.inflight[0]
            oldest.handle._collect()             # blocks until the oldest layer in flight is done
        return self.free.pop()

    def _release(self, slot: _Slot) -> None:
        with self.lock:
            if slot in self.inflight:
                self.inflight.remove(slot)
            self.free.append(slot)

    def _runner(self, slot: _Slot, p, m: int, n: int, h_kind: int, want_packed: bool, want_w_scaled: bool):
        key = (_params_signature(p), m, n, h_kind, want_packed, want_w_scaled, _lib.execution_mode())
        run = slot.runners.get(key)
        if run is not None:
            slot.runners.move_to_end(key)
            return run
        need = workspace_bytes(p, m, n, h_kind)
        if slot.ws is None or slot.ws.numel() < need:
            self._grow(slot, need)
        run = CalderaLayerRunner(p, m, n, h_kind, self.device, want_packed=want_packed, want_w_scaled=want_w_scaled,
                                 workspace=slot.ws)
        with torch.cuda.stream(slot.stream):
            run.capture()
        self._evict(self._runner_bytes(run), slot)
        self.bytes += self._runner_bytes(run)
        slot.runners[key] = run
        if run.small.numel() > slot.host_small.numel():
            slot.host_small = torch.empty(run.small.numel(), dtype=torch.float32).pin_memory()
        return run

    # ------------------------------------------------------------------ submission
    def submit(self, p, W: torch.Tensor, h_kind: int, H: Optional[torch.Tensor], seed: int, finish: Callable,
               want_packed: bool = True, want_w_scaled: bool = False,
               consume: Optional[Callable[[CalderaLayerRunner, dict], None]] = None) -> LayerHandle:
        """Enqueues one layer and returns immediately.  W / H: device tensors or pinned host tensors (staged with
        an asynchronous copy on the slot's stream).  `consume(run, kept)` runs right after the replay with the
        slot's stream current and must only enqueue device work (copies of the outputs to where they are going);
        tensors it stores in `kept` travel with the handle.  `finish(run, host_record, kept)` builds the result
        once the layer is done (host side, inside LayerHandle.result())."""
        m, n = int(W.shape[0]), int(W.shape[1])
        with self.lock, torch.cuda.device(self.device):
            slot = self._acquire()
            try:
                run = self._runner(slot, p, m, n, h_kind, want_packed, want_w_scaled)
                handle = LayerHandle(self, slot, run, finish)
                caller = torch.cuda.current_stream()
                slot.stream.wait_stream(caller)          # W / H may have been produced on the caller's stream
                with torch.cuda.stream(slot.stream):
                    run.launch(W, H, seed)
                    self.kernels_replayed += run.graph_kernels
                    if consume is not None:
                        consume(run, handle.kept)
                    slot.host_small[:run.small.numel()].copy_(run.small, non_blocking=True)
                    slot.event.record(slot.stream)
            except BaseException:
                self.free.append(slot)
                raise
            slot.handle = handle
            self.inflight.append(slot)
            return handle

    def drain(self) -> None:
        with self.lock:
            for slot in list(self.inflight):
                if slot.handle is not None:
                    slot.handle._collect()

    def release(self) -> None:
        """Drops every captured graph and arena (the engine stays usable)."""
        self.drain()
        with self.lock:
            for slot in self.slots:
                slot.runners.clear()
                slot.ws = None
            self.bytes = 0


_ENGINES = {}
_ENGINES_LOCK = threading.Lock()


def get_engine(device: torch.device, slots: Optional[int] = None) -> LayerEngine:
    """The per-device engine (created on first use; `slots` only takes effect then or when larger)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    with _ENGINES_LOCK:
        eng = _ENGINES.get(key)
        want = int(slots) if slots else int(os.environ.get("CB_ENGINE_SLOTS", "32"))
        if eng is None:
            eng = _ENGINES[key] = LayerEngine(torch.device("cuda", key), want)
        elif slots and want > len(eng.slots):
            with eng.lock, torch.cuda.device(eng.device):
                for i in range(len(eng.slots), want):
                    s = _Slot(eng.device, i)
                    eng.slots.append(s)
                    eng.free.insert(0, s)
        return eng


def release_engines() -> None:
    with _ENGINES_LOCK:
        for eng in _ENGINES.values():
            eng.release()
        _ENGINES.clear()
