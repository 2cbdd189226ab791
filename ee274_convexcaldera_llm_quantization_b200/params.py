"""Parameter / result containers, field-for-field as RCR/caldera/utils/dataclasses.py:11-113
(including the reference's defaults, e.g. update_order = [])."""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from .quantization import AbstractQuantizer, LowMemoryQuantizer, QuantizerFactory


@dataclass
class CalderaParams:
    """Parameters for the CALDERA decomposition (dataclasses.py:12-84)."""
    compute_quantized_component: bool = field(default=True)
    compute_low_rank_factors: bool = field(default=True)
    Q_bits: int = field(default=2)
    L_bits: int = field(default=2)
    R_bits: int = field(default=2)
    rank: int = field(default=64)
    iters: int = field(default=20)
    lplr_iters: int = field(default=5)
    activation_aware_LR: bool = field(default=True)
    update_order: list[str] = field(default_factory=list)
    quant_factory_Q: QuantizerFactory = field(default_factory=QuantizerFactory)
    quant_factory_LR: QuantizerFactory = field(default_factory=QuantizerFactory)
    rand_svd: bool = field(default=False)
    sigma_reg: float = field(default=0)


@dataclass
class CalderaDecomposition:
    """Components and parameters of the decomposition (dataclasses.py:88-106).

    Additive (B200 path only): Q_packed / L_packed / R_packed hold the bit-packed codes,
    best_step the sub-step index of the returned iterate."""
    Q: torch.Tensor = field(default=None)
    L: torch.Tensor = field(default=None)
    R: torch.Tensor = field(default=None)
    W: torch.Tensor = field(default=None)
    Q_idxs: torch.Tensor = field(default=None)
    L_idxs: torch.Tensor = field(default=None)
    R_idxs: torch.Tensor = field(default=None)
    Q_scale: float = field(default=1)
    L_scale: float = field(default=1)
    R_scale: float = field(default=1)
    global_scale: float = field(default=1)
    SU: torch.Tensor = field(default=None)
    SV: torch.Tensor = field(default=None)
    scaleWH: torch.Tensor = field(default=None)
    errors: dict[str, list[float]] = field(default_factory=dict)


@dataclass
class QuantInfo:
    """dataclasses.py:110-113."""
    quant: AbstractQuantizer = field(default_factory=LowMemoryQuantizer)
