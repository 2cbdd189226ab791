"""Model-level driver: the loop of the reference's `apply_CALDERA_quantization` (main.py:135-252) and its bit
accounting (main.py:326-335) on the asynchronous layer engine.

The reference walks `model.named_modules()`, decomposes one selected Linear layer at a time with a blocking
`caldera(..., device="cuda", scale_W=False)` call, rebuilds `Q + L @ R`, keeps the original weight when the relative
Frobenius error of the reconstruction exceeds a threshold, and counts quantised / unquantised parameters in module
globals.  Here the same selection rule, gate and counters are applied, but every selected layer is submitted to the
device's LayerEngine first (same-shape layers advance in batches, engine.py) and the reconstruction
`Q + L R` plus the two Frobenius norms are formed on the layer's stream right after its graph replay; the host only
reads two numbers per layer.  Nothing here is specific to one model family: any `torch.nn.Module` whose selected
sub-modules have a 2-D `.weight` works, and `hessians` maps module names to the diagonal of H (what main.py loads from
`diag_Hessians.pt`, :163-165).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch

from . import _lib
from .alg import caldera_async
from .engine import get_engine
from .hadamard import hadamard_transform
from .params import CalderaParams
from .quantization import QuantizerFactory
from .scheduler import batch_plan

# main.py:156
DEFAULT_LAYER_KEYS = ("mlp.up_proj", "mlp.down_proj", "mlp.gate_proj", "q_proj", "k_proj", "v_proj", "o_proj")


def default_quant_params() -> CalderaParams:
    """The parameters main.py:167-184 builds for every layer."""
    fac = QuantizerFactory(method="uniform", block_size=64)
    return CalderaParams(compute_quantized_component=True, compute_low_rank_factors=True, Q_bits=2, L_bits=16, R_bits=16,
                         rank=200, iters=5, lplr_iters=5, activation_aware_LR=True, update_order=["Q", "LR"],
                         quant_factory_Q=fac, quant_factory_LR=QuantizerFactory(method="uniform", block_size=64),
                         rand_svd=False, sigma_reg=1e-8)


@dataclass
class LayerRecord:
    name: str
    shape: tuple
    selected: bool
    applied: bool = False
    error: Optional[float] = None          # ||W - (Q + L R)||_F / ||W||_F (main.py:211)
    best_step: Optional[int] = None
    caldera_errors: Optional[dict] = None


@dataclass
class QuantizationReport:
    """The module globals of main.py (quantized_param_count, unquantized_language_param_count, vision_param_count)
    and the per-layer outcomes."""
    quantized_param_count: int = 0
    unquantized_language_param_count: int = 0
    vision_param_count: int = 0
    layers: List[LayerRecord] = field(default_factory=list)

    def bit_accounting(self, quantized_bits: int = 2, unquantized_bits: int = 4) -> dict:
        """main.py:326-335: total bits with the quantised layers at `quantized_bits` and the rest of the language model
        at `unquantized_bits`, against everything at `unquantized_bits`."""
        q, u = self.quantized_param_count, self.unquantized_language_param_count
        total = q * quantized_bits + u * unquantized_bits
        prior = (q + u) * unquantized_bits
        out = {"quantized_parameters": q, "unquantized_parameters": u, "total_bits": total, "prior_total_bits": prior}
        if q + u > 0:
            out["bits_ratio"] = total / prior
            out["quantized_fraction"] = q / (q + u)
        return out


def select_layers(model: torch.nn.Module, quantize_layer_list: Iterable[int], quantized_layer_limit: Optional[int] = None,
                  layer_keys: Sequence[str] = DEFAULT_LAYER_KEYS, name_filter: str = "language", min_size: int = 500):
    """The selection rule of main.py:147-163.  Returns [(name, module, kind)] in `named_modules()` order with kind
    "selected" (decompose), "language" (language-model weight that is left alone) or "other" (counted as vision)."""
    layer_ids = list(quantize_layer_list)
    out, taken = [], 0
    for name, module in model.named_modules():
        w = getattr(module, "weight", None)
        if w is None or not torch.is_tensor(w):
            continue
        if name_filter not in name:
            out.append((name, module, "other"))
            continue
        ok = (w.dim() == 2 and any(key in name for key in layer_keys) and w.size(0) > min_size and w.size(1) > min_size
              and any(f"layers.{i}" in name for i in layer_ids)
              and (quantized_layer_limit is None or taken < quantized_layer_limit))
        if ok:
            taken += 1
        out.append((name, module, "selected" if ok else "language"))
    return out


def _reconstruct_and_measure(lib, W: torch.Tensor, Q: torch.Tensor, L: torch.Tensor, R: torch.Tensor, gs: Optional[float]):
    """out = Q + L R (times global_scale when the layer was scaled) and device doubles (sum W, sum W^2, sum (W - out)^2),
    all enqueued on the current stream."""
    m, n = W.shape
    out = Q.clone()
    _lib.check(lib.cb_sgemm_strided(m, n, L.shape[1], 1.0, _lib.ptr(L), L.stride(0), L.stride(1), _lib.ptr(R), R.stride(0),
                                    R.stride(1), _lib.ptr(out), n, 1, 1, _lib.stream_ptr()), "sgemm")
    if gs is not None and gs != 1:
        out.mul_(gs)
    stats = torch.zeros(3, dtype=torch.float64, device=W.device)
    _lib.check(lib.cb_sum_stats(_lib.ptr(W), _lib.ptr(out), W.numel(), _lib.ptr(stats), _lib.stream_ptr()), "sum_stats")
    return out, stats


def apply_caldera_quantization(model: torch.nn.Module, hessians: Dict[str, torch.Tensor],
                               quant_params: Optional[CalderaParams] = None, *, device="cuda",
                               quantize_layer_list: Iterable[int] = range(0, 24), quantized_layer_limit: Optional[int] = None,
                               error_threshold: float = 1.0, hadamard: bool = False, scale_W: bool = False,
                               layer_keys: Sequence[str] = DEFAULT_LAYER_KEYS, name_filter: str = "language",
                               min_size: int = 500, slots: Optional[int] = None, batch: Optional[int] = None,
                               seed: int = 0, on_layer: Optional[Callable[[LayerRecord], None]] = None) -> QuantizationReport:
    """Replaces the weights of the selected layers by `Q + L R` in place (main.py:189-217) and returns the counters.

    hessians: module name -> diagonal of H (fp32-convertible, length in_features); a missing name raises KeyError like
    `Hall[name]` (main.py:163).  scale_W defaults to False as in main.py:195.  `hadamard=True` decomposes
    H1 W H2 instead (main.py:218-240; both dimensions must be powers of two, since the reference's zero padding changes
    the number of columns H has to match).  `on_layer(record)` is called as each layer's outcome becomes known."""
    qp = quant_params if quant_params is not None else default_quant_params()
    dev = torch.device(device) if not isinstance(device, torch.device) else device
    if dev.type != "cuda":
        raise RuntimeError("apply_caldera_quantization: a CUDA device is required (there is no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    report = QuantizationReport()
    plan = select_layers(model, quantize_layer_list, quantized_layer_limit, layer_keys, name_filter, min_size)
    engine = get_engine(dev, slots, batch)
    # same-shape layers are split into equal groups up front (a captured batch always advances all of its layers)
    per_shape: Dict[tuple, List[str]] = {}
    for name, module, kind in plan:
        if kind == "selected":
            per_shape.setdefault(tuple(module.weight.shape), []).append(name)
    group_of: Dict[str, int] = {}
    for names in per_shape.values():
        pos = 0
        for size in batch_plan(len(names), engine.batch):
            for nm in names[pos:pos + size]:
                group_of[nm] = size
            pos += size
    pending = []
    with torch.no_grad(), torch.cuda.device(dev):
        for name, module, kind in plan:
            w = module.weight
            rec = LayerRecord(name=name, shape=tuple(w.shape), selected=kind == "selected")
            report.layers.append(rec)
            if kind == "other":
                report.vision_param_count += w.numel()
                continue
            if kind == "language":
                report.unquantized_language_param_count += w.numel()
                continue
            h = hessians[name].to(dev, torch.float32)
            Wd = w.data.to(dev, torch.float32)
            if hadamard:
                m, n = Wd.shape
                if (m & (m - 1)) or (n & (n - 1)):
                    raise ValueError(f"hadamard=True needs power-of-two dimensions, {name} is {m} x {n}")
                Wd, _ = hadamard_transform(Wd, inverse=False)

            def consume(run, kept, Wd=Wd):
                out, stats = _reconstruct_and_measure(lib, Wd, run.Q, run.L, run.R, None)
                kept["out"], kept["stats"] = out, stats

            hd = caldera_async(qp, Wd, h, device=dev, use_tqdm=False, scale_W=scale_W, W_copy="none", seed=seed,
                               return_dense=True, return_packed=False, consume=consume, slots=slots, batch=batch,
                               batch_hint=group_of[name])
            pending.append((rec, module, Wd, hd))
        engine.flush()
        for rec, module, Wd, hd in pending:
            dec = hd.result()
            out, stats = hd.kept["out"], hd.kept["stats"]
            gs = float(dec.global_scale) if scale_W else 1.0
            if gs != 1.0:
                # caldera() returns Q, L, R in the scaled space (alg.py:42): rebuild in the caller's units
                out, stats = _reconstruct_and_measure(lib, Wd, dec.Q, dec.L, dec.R, gs)
            _, w_sq, diff_sq = stats.tolist()
            rec.error = (diff_sq / w_sq) ** 0.5 if w_sq > 0 else 0.0
            rec.best_step, rec.caldera_errors = dec.best_step, dec.errors
            if rec.error > error_threshold:                      # main.py:213-216
                report.unquantized_language_param_count += module.weight.numel()
            else:
                if hadamard:
                    out = hadamard_transform(out, inverse=True, original_shape=tuple(module.weight.shape))
                module.weight.data = out.to(module.weight.dtype).to(module.weight.device)
                rec.applied = True
                report.quantized_param_count += module.weight.numel()
            if on_layer is not None:
                on_layer(rec)
    return report
