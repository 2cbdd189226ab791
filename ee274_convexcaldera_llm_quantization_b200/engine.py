"""Asynchronous layer engine: many decompositions in flight on one GPU, driven by ONE host thread.

The reference decomposes layers one after the other and blocks on every one (main.py:147-199; host syncs at
alg.py:39, :59, :300).  Here a layer is part of a CUDA-graph replay whose only latency-bound part is a chain of
single-CTA factorisation kernels, so throughput comes from keeping many independent layers in flight.  Round 1 did
that with one blocked host thread per layer, which made the end-to-end rate `threads / layer latency` and capped it
by the host's core count.  The engine decouples the two:

  submit()   stages W and h (asynchronous H2D copies when they are pinned host tensors) into a free slot and returns
             a LayerHandle at once.  Layers that the batched driver supports (csrc/driver.cu: cb_caldera_batch) are
             collected until a slot's batch is full -- `batch` same-shape layers then advance in lock step as ONE
             graph replay, each contraction / factorisation being one launch for all of them -- other configurations
             replay a single-layer graph.  After the replay the caller's device-side `consume` copies (e.g. straight
             into a wire-format arena or pinned host buffers) and the D2H copy of the ~100-byte result records are
             enqueued and an event is recorded.  submit() only blocks when every slot is busy.
  LayerHandle.result()  launches the handle's batch if it is still filling, waits for its event and builds the
             CalderaDecomposition.

Graph launches go through one launcher thread per slot: cudaGraphLaunch of a several-hundred-node graph blocks the
calling thread for tens of milliseconds whenever the device's queues are deep (measured: 2-10 ms typically, up to
100 ms, scripts/probe_e2e_trace.py), and a submitting thread that is stuck there cannot stage the next batch's
input copies -- the host link then idles and the end-to-end rate drops by up to 2x.  The caller's thread only stages
inputs and hands full batches over; the slot's launcher replays, runs the `consume` hooks and records the slot's
event (one thread per slot, so that a launch that blocks does not hold up the launches of the other slots).

Host inputs of every slot cross the link on ONE engine-wide copy stream, in submission order.  Copies enqueued on the
slots' own streams are served side by side: when six slots are filled at once (the start of a job) all of them receive
their last layer at about the same time, the first graph starts after the whole 9 GB instead of after its own 1.5 GB,
and the slots then run -- and ask for their next inputs -- in lock step.  In FIFO order slot 0 starts after 28 ms and the
slots stay staggered.  CB_ENGINE_COPY_FIFO=0 restores per-slot copies.

The first engine of a process also calls gc.freeze() (see _freeze_import_time_objects: full garbage collections of
torch's import-time objects were the other source of 0.1-0.3 s stalls).

Slots own a stream and one device arena that the graphs of every shape captured in that slot share (they are never in
flight together), so device memory is slots x (largest batch footprint) whatever the number of distinct shapes; the
arena is bounded by `CB_ENGINE_MAX_BYTES` (default 60 % of the device memory free at creation) through the batch size.
"""
from __future__ import annotations

import ctypes as C
import os
import queue
import threading
from typing import Callable, Dict, List, Optional

import torch

from . import _lib
from .runner import BatchRunner, CalderaLayerRunner, workspace_bytes


def _params_signature(p) -> tuple:
    return tuple(getattr(p, name) if name != "order" else tuple(p.order)
                 for name, _ in p._fields_ if name != "seed")


class LayerHandle:
    """One layer in flight.  `result()` blocks until it is done and returns what `finish` builds."""

    def __init__(self, group: "_Group", index: int, finish: Callable, consume: Optional[Callable], seed: int):
        self._group, self._index, self._finish, self._consume, self._seed = group, index, finish, consume, seed
        self._value = None
        self._done = False
        self.kept: dict = {}           # device tensors produced by `consume` that belong to this layer

    @property
    def run(self):
        return self._group.views[self._index]

    def done(self) -> bool:
        g = self._group
        return g.host is not None or (g.launched and g.issued.is_set() and g.error is None and g.slot.event.query())

    def result(self):
        if not self._done:
            g = self._group
            g.engine._collect(g)
            self._value = self._finish(g.views[self._index], g.host[self._index], self.kept)
            self._done = True
            self._finish = self._consume = None
        return self._value


class _ViewMeta:
    """What is left of a layer's view once its group was collected: the layout of its host record."""
    __slots__ = ("nsteps", "nerr_pad")

    def __init__(self, nsteps: int, nerr_pad: int):
        self.nsteps, self.nerr_pad = nsteps, nerr_pad


class _Group:
    """The layers sharing one graph replay in one slot (a full or partial batch, or a single layer)."""

    def __init__(self, engine, slot, runner, views, key):
        self.engine, self.slot, self.runner, self.views, self.key = engine, slot, runner, views, key
        self.handles: List[LayerHandle] = []
        self.launched = False                    # handed to the launcher thread
        self.issued = threading.Event()          # the launcher has enqueued everything on the slot's stream
        self.copy_done: Optional[torch.cuda.Event] = None    # host inputs staged on the engine's copy stream
        self.fifo_inputs = False
        self.error: Optional[BaseException] = None
        self.host = None


class _Slot:
    def __init__(self, device: torch.device, index: int):
        self.index = index
        self.stream = torch.cuda.Stream(device=device)
        # blocking: the waiting thread sleeps instead of spinning (8 ranks share the box's host cores)
        self.event = torch.cuda.Event(blocking=os.environ.get("CB_ENGINE_SPIN_WAIT", "0") != "1")
        self.arena: Optional[torch.Tensor] = None
        self.runners: Dict[tuple, object] = {}
        self.host_small: Optional[torch.Tensor] = None
        self.group: Optional[_Group] = None
        self.work: "queue.SimpleQueue[Optional[_Group]]" = queue.SimpleQueue()
        self.launcher: Optional[threading.Thread] = None


class LayerEngine:
    def __init__(self, device: torch.device, slots: int = 3, batch: int = 16, max_bytes: Optional[int] = None):
        self.device = device
        self.lock = threading.RLock()
        self.batch = max(1, int(batch))
        with torch.cuda.device(device):
            self.slots = [_Slot(device, i) for i in range(max(1, int(slots)))]
            if max_bytes is None:
                env = os.environ.get("CB_ENGINE_MAX_BYTES")
                max_bytes = int(env) if env else int(0.6 * torch.cuda.mem_get_info(device)[0])
        self.max_bytes = int(max_bytes)
        with torch.cuda.device(device):
            self.copy_stream = torch.cuda.Stream(device=device) if os.environ.get("CB_ENGINE_COPY_FIFO", "1") != "0" else None
        self.free = list(reversed(self.slots))       # pop() hands out slot 0 first
        self.inflight: List[_Slot] = []              # launched, in launch order
        self.filling: Dict[tuple, _Group] = {}       # key -> group still collecting layers
        self.captures = 0                            # graphs captured so far (a warm job captures none)

    # ------------------------------------------------------------------ slots
    def _slot_bytes(self) -> int:
        return self.max_bytes // len(self.slots)

    def _acquire(self, key=None) -> _Slot:
        if not self.free:
            if self.inflight:
                self._collect(self.inflight[0].group)           # blocks until the oldest group in flight is done
            else:
                # every slot is filling a different shape: launch the fullest partial batch to make room
                g = max(self.filling.values(), key=lambda g_: len(g_.handles))
                self._launch(g)
                self._collect(g)
        # plain LIFO (slot 0 first on an idle engine, see _collect): a job that is repeated on an idle engine is dealt
        # the same slot sequence again and therefore meets the graphs its first pass captured
        return self.free.pop()

    def _arena(self, slot: _Slot, nbytes: int) -> torch.Tensor:
        if slot.arena is None or slot.arena.numel() < nbytes:
            slot.runners.clear()                                 # graphs captured against the old arena die with it
            slot.arena = None
            slot.arena = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return slot.arena

    def reserve(self, nbytes: int) -> None:
        """Sizes every slot's arena for the largest batch of a job up front (graphs captured later share it)."""
        with self.lock, torch.cuda.device(self.device):
            self.drain()
            for slot in self.slots:
                self._arena(slot, min(nbytes, self._slot_bytes()))

    def batch_for(self, p, m: int, n: int, h_kind: int, want_packed: bool, want_w_scaled: bool, hint: Optional[int],
                  want_dense: bool = True) -> int:
        """Layers per graph replay for this configuration: 0 = not batchable (single-layer graph)."""
        lib = _lib.load()
        if self.batch <= 1 or not lib.cb_caldera_batch_supported(C.byref(p), m, n, h_kind):
            return 0
        b = self.batch if not hint else max(1, min(self.batch, int(hint)))
        stride = BatchRunner.slab_stride(p, m, n, h_kind, want_packed, want_w_scaled, want_dense)
        return max(1, min(b, self._slot_bytes() // stride))

    def _group_for(self, p, m, n, h_kind, want_packed, want_w_scaled, hint, want_dense=True) -> _Group:
        nb = self.batch_for(p, m, n, h_kind, want_packed, want_w_scaled, hint, want_dense)
        if nb == 0:
            want_dense = True                 # the single-layer driver always materialises Q
        key = (_params_signature(p), m, n, h_kind, want_packed, want_w_scaled, want_dense, nb)
        g = self.filling.get(key)
        if g is not None:
            return g
        slot = self._acquire(key)
        try:
            run = slot.runners.get(key)
            if run is None:
                self.captures += 1
                with torch.cuda.stream(slot.stream):
                    if nb > 0:
                        stride = BatchRunner.slab_stride(p, m, n, h_kind, want_packed, want_w_scaled, want_dense)
                        run = BatchRunner(p, m, n, h_kind, nb, self.device, want_packed=want_packed, want_w_scaled=want_w_scaled,
                                          want_dense=want_dense, slab=self._arena(slot, max(stride * nb, slot.arena.numel() if slot.arena is not None else 0)))
                    else:
                        need = workspace_bytes(p, m, n, h_kind)
                        run = CalderaLayerRunner(p, m, n, h_kind, self.device, want_packed=want_packed, want_w_scaled=want_w_scaled,
                                                 workspace=self._arena(slot, max(need, slot.arena.numel() if slot.arena is not None else 0)))
                    run.capture()
                slot.runners[key] = run
            views = run.layers if nb > 0 else [run]
            nsmall = views[0].small.numel()
            if slot.host_small is None or slot.host_small.numel() < len(views) * nsmall:
                slot.host_small = torch.empty(len(views) * nsmall, dtype=torch.float32).pin_memory()
        except BaseException:
            self.free.append(slot)
            raise
        g = _Group(self, slot, run, views, key)
        slot.group = g
        self.filling[key] = g
        return g

    # ------------------------------------------------------------------ submission
    def submit(self, p, W: torch.Tensor, h_kind: int, H: Optional[torch.Tensor], seed: int, finish: Callable,
               want_packed: bool = True, want_w_scaled: bool = False,
               consume: Optional[Callable] = None, batch_hint: Optional[int] = None, want_dense: bool = True) -> LayerHandle:
        """Stages one layer and returns immediately.  W / H: device tensors or pinned host tensors.  `consume(view,
        kept)` runs right after the layer's graph replay was enqueued, with the slot's stream current, and must only
        enqueue device work (copies of view.Q_packed, view.L, ... to where they are going); tensors it stores in
        `kept` travel with the handle.  `finish(view, host_record, kept)` builds the result once the layer is done
        (host side, inside LayerHandle.result()).  `batch_hint`: how many layers of this shape the caller is about
        to submit (bounds the batch size so that a short job does not run a mostly empty batch)."""
        m, n = int(W.shape[0]), int(W.shape[1])
        with self.lock, torch.cuda.device(self.device):
            g = self._group_for(p, m, n, h_kind, want_packed, want_w_scaled, batch_hint, want_dense)
            idx = len(g.handles)
            handle = LayerHandle(g, idx, finish, consume, seed)
            caller = torch.cuda.current_stream()
            # host inputs: the engine-wide copy stream (FIFO over all slots, see the module docstring); the slot is free,
            # i.e. its previous group was collected, so nothing on its own stream still reads the buffers written here
            fifo = self.copy_stream is not None and not W.is_cuda
            st = self.copy_stream if fifo else g.slot.stream
            g.fifo_inputs = g.fifo_inputs or fifo
            st.wait_stream(caller)                         # W / H may have been produced on the caller's stream
            with torch.cuda.stream(st):
                if isinstance(g.runner, BatchRunner):
                    g.runner.stage(idx, W, H)
                else:
                    g.runner.W_in.copy_(W, non_blocking=True)
                    if g.runner.h_in is not None and H is not None:
                        g.runner.h_in.copy_(H, non_blocking=True)
            g.handles.append(handle)
            if len(g.handles) == len(g.views):
                self._launch(g)
            return handle

    def _launch(self, g: _Group) -> None:
        """Hands a group to the launcher thread (idempotent).  Called with the engine lock held."""
        if g.launched:
            return
        g.launched = True
        self.filling.pop(g.key, None)
        if g.fifo_inputs:
            g.copy_done = torch.cuda.Event()
            g.copy_done.record(self.copy_stream)
        self.inflight.append(g.slot)
        slot = g.slot
        if slot.launcher is None or not slot.launcher.is_alive():
            slot.launcher = threading.Thread(target=self._launcher_loop, args=(slot,), daemon=True,
                                             name=f"caldera-launcher-{self.device.index}-{slot.index}")
            slot.launcher.start()
        slot.work.put(g)

    def _launcher_loop(self, slot: "_Slot") -> None:
        # never takes the engine lock: a group in the queue is owned by this thread until `issued` is set
        while True:
            g = slot.work.get()
            if g is None:
                return
            try:
                self._issue(g)
            except BaseException as exc:      # noqa: BLE001  (re-raised in the caller's thread by _collect)
                g.error = exc
            finally:
                g.issued.set()

    def _issue(self, g: _Group) -> None:
        slot = g.slot
        with torch.cuda.device(self.device), torch.cuda.stream(slot.stream):
            if g.copy_done is not None:
                slot.stream.wait_event(g.copy_done)
            seeds = [h._seed for h in g.handles]
            if isinstance(g.runner, BatchRunner):
                g.runner.replay(seeds)
            else:
                g.runner.seed_dev.fill_(int(seeds[0]))
                g.runner.graph.replay()
                g.runner.lib.cb_note_launches(g.runner.graph_kernels)
            for h in g.handles:
                if h._consume is not None:
                    h._consume(g.views[h._index], h.kept)
            nsmall = g.views[0].small.numel()
            dst = slot.host_small[:len(g.views) * nsmall].view(len(g.views), nsmall)
            if isinstance(g.runner, BatchRunner):
                dst.copy_(g.runner.small_all, non_blocking=True)
            else:
                dst[0].copy_(g.runner.small, non_blocking=True)
            slot.event.record(slot.stream)

    def _collect(self, g: _Group) -> None:
        """Waits for a group and frees its slot (idempotent)."""
        with self.lock:
            if g.host is not None:
                return
            self._launch(g)
            slot = g.slot
            g.issued.wait()
            if g.error is None:
                slot.event.synchronize()
            nsmall = g.views[0].small.numel()
            g.host = slot.host_small[:len(g.views) * nsmall].view(len(g.views), nsmall).clone()
            # a finished group must not keep the slot's arena alive (the caller may hold its handles for a long time):
            # `finish` only needs the record layout
            g.views = [_ViewMeta(v.nsteps, v.nerr_pad) for v in g.views]
            g.runner = None
            # handle -> group only from here on, so handles, groups and the device tensors a handle keeps die by
            # reference counting.  With the cycle in place they were left to the cyclic collector, whose generation-2
            # passes then had real work to do in the middle of a job: one 0.3-0.55 s step in most 10-step runs
            # (e2e 460-640 matrices/s; 669-673 in five consecutive runs without it)
            g.handles = []
            slot.group = None
            if slot in self.inflight:
                self.inflight.remove(slot)
            self.free.append(slot)
            if len(self.free) == len(self.slots):
                # idle: hand the slots out in the same order as a fresh engine, so that a repeated job meets its graphs
                self.free.sort(key=lambda s_: -s_.index)
            if g.error is not None:
                raise g.error

    def flush(self) -> None:
        """Launches every partially filled batch."""
        with self.lock:
            for g in list(self.filling.values()):
                self._launch(g)

    def drain(self) -> None:
        with self.lock:
            self.flush()
            for slot in list(self.inflight):
                if slot.group is not None:
                    self._collect(slot.group)

    def release(self) -> None:
        """Drops every captured graph and arena (the engine stays usable)."""
        self.drain()
        with self.lock:
            for slot in self.slots:
                slot.runners.clear()
                slot.arena = None
            for slot in self.slots:
                if slot.launcher is not None:
                    slot.work.put(None)                  # restarted by the slot's next launch
                    slot.launcher.join(timeout=10.0)
                    slot.launcher = None


_ENGINES: Dict[int, LayerEngine] = {}
_ENGINES_LOCK = threading.Lock()
_GC_FROZEN = False


def _freeze_import_time_objects() -> None:
    """A process that has imported torch holds a few million long-lived Python objects, and a full (generation 2)
    garbage collection walks all of them: 0.1-0.3 s, during which neither the submitting thread nor a slot's launcher
    runs.  At ~650 layers/s the engine's per-layer handles, closures and records trigger one about every second
    (measured: scripts/probe_e2e_trace.py, profiles/r2_e2e_trace.md -- a launcher stuck for 330 ms in its consume hooks,
    end-to-end rate 560 instead of 660 matrices/s).  gc.freeze() moves what exists now to the permanent generation, so
    later collections only look at what was allocated since.  Opt out with CB_ENGINE_GC_FREEZE=0."""
    global _GC_FROZEN
    if _GC_FROZEN or os.environ.get("CB_ENGINE_GC_FREEZE", "1") == "0":
        return
    import gc
    gc.collect()
    gc.freeze()
    _GC_FROZEN = True


def get_engine(device: torch.device, slots: Optional[int] = None, batch: Optional[int] = None) -> LayerEngine:
    """The per-device engine, created on first use: `slots` groups in flight (default CB_ENGINE_SLOTS or 3) of up to
    `batch` layers each (default CB_ENGINE_BATCH or 16).  Asking for a different geometry later rebuilds it."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    with _ENGINES_LOCK:
        eng = _ENGINES.get(key)
        want_slots = int(slots) if slots else int(os.environ.get("CB_ENGINE_SLOTS", "3"))
        want_batch = int(batch) if batch else int(os.environ.get("CB_ENGINE_BATCH", "16"))
        if eng is not None and ((slots and want_slots != len(eng.slots)) or (batch and want_batch != eng.batch)):
            eng.release()
            eng = None
        if eng is None:
            _freeze_import_time_objects()
            eng = _ENGINES[key] = LayerEngine(torch.device("cuda", key), want_slots, want_batch)
        return eng


def release_engines() -> None:
    with _ENGINES_LOCK:
        for eng in _ENGINES.values():
            eng.release()
        _ENGINES.clear()
