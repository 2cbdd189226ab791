"""The model-level driver (main.py:135-252, :326-335): layer selection, error gate, counters, bit accounting."""
import numpy as np
import pytest
import torch
from torch import nn

from ee274_convexcaldera_llm_quantization_b200.model_driver import (QuantizationReport, apply_caldera_quantization,
                                                                     select_layers)
from src.caldera.utils.dataclasses import CalderaParams
from src.caldera.utils.quantization import QuantizerFactory


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.o_proj = (nn.Linear(d, d, bias=False) for _ in range(4))


class _Mlp(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.gate_proj, self.up_proj, self.down_proj = nn.Linear(d, f, bias=False), nn.Linear(d, f, bias=False), nn.Linear(f, d, bias=False)


class _Block(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.self_attn, self.mlp, self.input_layernorm = _Attn(d), _Mlp(d, f), nn.LayerNorm(d)


class _Toy(nn.Module):
    def __init__(self, d=512, f=1024, blocks=3):
        super().__init__()
        self.language_model = nn.ModuleDict({"layers": nn.ModuleList([_Block(d, f) for _ in range(blocks)]),
                                             "lm_head": nn.Linear(d, 600, bias=False), "small": nn.Linear(d, 400, bias=False)})
        self.vision_tower = nn.ModuleDict({"fc1": nn.Linear(64, 96)})


def _toy(seed=0):
    torch.manual_seed(seed)
    model = _Toy()
    for p in model.parameters():
        if p.dim() == 2:
            p.data.normal_(0, 0.02)
    return model


def test_select_layers_follows_the_reference_rule():
    model = _toy()
    plan = select_layers(model, quantize_layer_list=[0, 2])
    kinds = {name: kind for name, _, kind in plan}
    assert kinds["language_model.layers.0.self_attn.q_proj"] == "selected"
    assert kinds["language_model.layers.2.mlp.down_proj"] == "selected"
    assert kinds["language_model.layers.1.mlp.up_proj"] == "language"         # layer 1 is not in the list
    assert kinds["language_model.layers.0.input_layernorm"] == "language"     # has a weight, is not a projection
    assert kinds["language_model.lm_head"] == "language"                      # not one of the projection names
    assert kinds["vision_tower.fc1"] == "other"
    assert sum(k == "selected" for k in kinds.values()) == 14
    limited = select_layers(model, quantize_layer_list=[0, 1, 2], quantized_layer_limit=5)
    assert sum(kind == "selected" for _, _, kind in limited) == 5
    big_only = select_layers(model, quantize_layer_list=[0], min_size=600)
    assert [n for n, _, k in big_only if k == "selected"] == []               # 512-wide layers fail size > 600


def test_bit_accounting():
    rep = QuantizationReport(quantized_param_count=300, unquantized_language_param_count=100)
    acc = rep.bit_accounting()
    assert acc["total_bits"] == 300 * 2 + 100 * 4 and acc["prior_total_bits"] == 400 * 4
    assert acc["bits_ratio"] == pytest.approx(1000 / 1600) and acc["quantized_fraction"] == pytest.approx(0.75)
    assert "bits_ratio" not in QuantizationReport().bit_accounting()


def _params():
    fac = QuantizerFactory(method="uniform", block_size=64)
    return CalderaParams(Q_bits=2, L_bits=16, R_bits=16, rank=32, iters=2, lplr_iters=2, activation_aware_LR=True,
                         update_order=["Q", "LR"], quant_factory_Q=fac, quant_factory_LR=fac, rand_svd=False, sigma_reg=1e-8)


@pytest.mark.gpu
def test_apply_replaces_selected_weights_and_counts():
    from src.caldera.decomposition.alg import caldera
    model = _toy(1)
    before = {n: m.weight.detach().clone() for n, m in model.named_modules() if hasattr(m, "weight")}
    g = torch.Generator().manual_seed(5)
    hess = {n: 0.5 + torch.rand(m.weight.shape[1], generator=g) for n, m in model.named_modules()
            if isinstance(m, nn.Linear) and "language" in n}
    seen = []
    rep = apply_caldera_quantization(model, hess, _params(), device="cuda", quantize_layer_list=[0, 2], error_threshold=1.0,
                                     on_layer=seen.append)
    selected = [r for r in rep.layers if r.selected]
    assert len(selected) == 14 and all(r.applied for r in selected) and len(seen) == 14
    numel = lambda pred: sum(before[n].numel() for n in before if pred(n))            # noqa: E731
    sel_names = {r.name for r in selected}
    assert rep.quantized_param_count == numel(lambda n: n in sel_names)
    assert rep.unquantized_language_param_count == numel(lambda n: "language" in n and n not in sel_names)
    assert rep.vision_param_count == before["vision_tower.fc1"].numel()
    acc = rep.bit_accounting()
    assert acc["total_bits"] == rep.quantized_param_count * 2 + rep.unquantized_language_param_count * 4
    for n, m in model.named_modules():
        if not hasattr(m, "weight"):
            continue
        if n in sel_names:
            assert not torch.equal(m.weight.data.cpu(), before[n])
        else:
            assert torch.equal(m.weight.data.cpu(), before[n])
    # one layer against a blocking caldera() call with the same arguments (main.py:189-197)
    name = "language_model.layers.2.mlp.up_proj"
    rec = next(r for r in selected if r.name == name)
    dec = caldera(_params(), before[name], torch.diag_embed(hess[name]), device="cuda", use_tqdm=False, scale_W=False)
    out = (dec.Q + dec.L @ dec.R).cpu()
    err = float(torch.norm(before[name] - out) / torch.norm(before[name]))
    assert rec.error == pytest.approx(err, rel=2e-3)
    got = dict(model.named_modules())[name].weight.data.cpu()
    assert float(torch.norm(before[name] - got) / torch.norm(before[name])) == pytest.approx(rec.error, rel=1e-5)
    assert rec.caldera_errors["LR"][-1] == pytest.approx(dec.errors["LR"][-1], rel=3e-3)


@pytest.mark.gpu
def test_error_gate_keeps_the_original_weight():
    model = _toy(2)
    before = {n: m.weight.detach().clone() for n, m in model.named_modules() if hasattr(m, "weight")}
    hess = {n: torch.ones(m.weight.shape[1]) for n, m in model.named_modules() if isinstance(m, nn.Linear)}
    rep = apply_caldera_quantization(model, hess, _params(), device="cuda", quantize_layer_list=[1], error_threshold=1e-6)
    assert rep.quantized_param_count == 0 and not any(r.applied for r in rep.layers)
    assert all(r.error > 1e-6 for r in rep.layers if r.selected)
    for n, m in model.named_modules():
        if hasattr(m, "weight"):
            assert torch.equal(m.weight.data.cpu(), before[n])
    total_language = sum(v.numel() for n, v in before.items() if "language" in n)
    assert rep.unquantized_language_param_count == total_language
    with pytest.raises(KeyError):
        apply_caldera_quantization(model, {}, _params(), device="cuda", quantize_layer_list=[1])


@pytest.mark.gpu
def test_scaled_and_hadamard_variants():
    model = _toy(3)
    hess = {n: torch.ones(m.weight.shape[1]) for n, m in model.named_modules() if isinstance(m, nn.Linear)}
    w0 = model.language_model["layers"][0].mlp.up_proj.weight.detach().clone()
    rep = apply_caldera_quantization(model, hess, _params(), device="cuda", quantize_layer_list=[0], quantized_layer_limit=7,
                                     scale_W=True)
    rec = next(r for r in rep.layers if r.name.endswith("layers.0.mlp.up_proj"))
    got = model.language_model["layers"][0].mlp.up_proj.weight.data.cpu()
    assert float(torch.norm(w0 - got) / torch.norm(w0)) == pytest.approx(rec.error, rel=1e-5) and 0.1 < rec.error < 1.0
    model = _toy(3)
    rep_h = apply_caldera_quantization(model, hess, _params(), device="cuda", quantize_layer_list=[0], quantized_layer_limit=7,
                                       hadamard=True)
    rec_h = next(r for r in rep_h.layers if r.name.endswith("layers.0.mlp.up_proj"))
    got_h = model.language_model["layers"][0].mlp.up_proj.weight.data.cpu()
    # the transform is orthogonal: the error measured in the rotated space is the error of the recovered weight
    assert float(torch.norm(w0 - got_h) / torch.norm(w0)) == pytest.approx(rec_h.error, rel=1e-4)
