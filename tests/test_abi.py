"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol that
include/caldera_b200.h declares; the host shim mirrors the reference's names, defaults and
error conventions (RCR/caldera/utils/quantization.py:23-55, 246-255; dataclasses.py:11-113)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from ee274_convexcaldera_llm_quantization_b200 import _lib, build as cbuild


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        cbuild.build()
    return _lib.load()


def declared_symbols(measure=False):
    text = open(os.path.join(ROOT, "include", "caldera_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    release, _, aids = text.partition("#ifdef CB_MEASURE")
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", aids if measure else release)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/caldera_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.exported_symbols()) == names


def test_release_library_has_no_process_wide_state(lib):
    """re-entrant, no globals.  The execution policy is a field of cb_caldera_params; the measurement
    aids (timing stamps, probes, knock-out switch) are declared under CB_MEASURE and compiled only into
    libcaldera_b200_measure.so."""
    raw = ctypes.CDLL(_lib.LIB_PATH)
    aids = declared_symbols(measure=True)
    assert "cb_set_gemm_timing" in aids and "cb_probe_mma_rate" in aids
    for n in aids + ["cb_set_execution_mode", "cb_set_gemm_target_ctas"]:
        assert not hasattr(raw, n), f"{n} must not be exported by the release library"
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = [ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("cb_")]
    assert not [n for n in exported if n.startswith("cb_set_") or n.startswith("cb_probe_")], exported
    assert "exec_mode" in [f for f, _ in _lib.cb_caldera_params._fields_]


def test_version_and_status_strings(lib):
    assert lib.cb_version() == 100
    assert _lib.status_string(0) == "ok"
    assert "Bit-width" in _lib.status_string(_lib.CB_ERR_BITS)
    assert lib.cb_packed_bytes(17, 2) == 5 and lib.cb_packed_bytes(17, 4) == 9
    assert lib.cb_packed_bytes(17, 8) == 17 and lib.cb_packed_bytes(17, 16) == 34


def test_params_struct_matches_header(lib):
    # sizeof(cb_caldera_params) as the C compiler lays it out: 10 int32 + int32[8] + ... (see header)
    p = _lib.cb_caldera_params()
    assert ctypes.sizeof(p) % 8 == 0
    # workspace query is pure host code: usable without a GPU
    p.compute_q = p.compute_lr = 1
    p.q_bits, p.l_bits, p.r_bits = 2, 16, 16
    p.rank, p.iters, p.lplr_iters, p.aware, p.n_order = 128, 5, 5, 1, 2
    p.order[0], p.order[1] = 0, 1
    p.scale_w, p.power_iters = 1, -1
    b = lib.cb_caldera_layer_workspace_bytes(ctypes.byref(p), 4096, 4096, _lib.CB_H_DIAG)
    assert 200e6 < b < 600e6
    p.rank = 5000
    assert lib.cb_caldera_layer_workspace_bytes(ctypes.byref(p), 4096, 4096, _lib.CB_H_DIAG) == 0


def test_reference_import_paths_and_defaults():
    from src.caldera.utils.dataclasses import CalderaParams, CalderaDecomposition, QuantInfo
    from src.caldera.utils.quantization import QuantizerFactory, LowMemoryQuantizer
    from src.caldera.decomposition.alg import caldera  # noqa: F401
    p = CalderaParams()
    assert (p.Q_bits, p.L_bits, p.R_bits, p.rank, p.iters, p.lplr_iters) == (2, 2, 2, 64, 20, 5)
    assert p.update_order == [] and p.rand_svd is False and p.sigma_reg == 0
    assert p.activation_aware_LR is True and p.compute_low_rank_factors and p.compute_quantized_component
    assert isinstance(p.quant_factory_Q, QuantizerFactory) and p.quant_factory_Q.block_size == 64
    d = CalderaDecomposition()
    assert d.Q is None and d.Q_scale == 1 and d.global_scale == 1 and d.errors == {}
    assert isinstance(QuantInfo().quant, LowMemoryQuantizer)
    assert str(QuantizerFactory()) == "QuantizerFactory(method=uniform, block_size=64)"


def test_quantizer_error_conventions():
    from src.caldera.utils.quantization import QuantizerFactory, LowMemoryQuantizer
    with pytest.raises(AssertionError):
        LowMemoryQuantizer(num_bits=3)
    with pytest.raises(NotImplementedError):
        LowMemoryQuantizer(num_bits=4, method="lattice")
    with pytest.raises(ValueError):
        LowMemoryQuantizer(num_bits=2, method="nf4")
    with pytest.raises(ValueError):
        LowMemoryQuantizer(num_bits=4, method="nf2")
    with pytest.raises(ValueError):
        LowMemoryQuantizer(num_bits=2, method="bbint4")
    q = QuantizerFactory(method="uniform", block_size=64).get_quantizer(4)
    with pytest.raises(ValueError):
        q.quantize_block(torch.zeros(2, 3, 4))
    with pytest.raises(ValueError):
        q.quantize_block(torch.zeros(3, 5))
    # no CPU fallback: a CPU tensor is rejected loudly
    with pytest.raises(RuntimeError, match="no CPU"):
        q.quantize_block(torch.zeros(4, 64))


def test_caldera_rejects_cpu_device():
    from src.caldera.utils.dataclasses import CalderaParams
    from src.caldera.decomposition.alg import caldera
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        caldera(CalderaParams(update_order=["Q", "LR"]), torch.zeros(8, 8), device="cpu")
    with pytest.raises(ValueError):
        caldera(CalderaParams(update_order=["Q", "LR"]), torch.zeros(8), device="cpu")


def test_make_c_params_roundtrip():
    from ee274_convexcaldera_llm_quantization_b200.alg import make_c_params
    from src.caldera.utils.dataclasses import CalderaParams
    cp = make_c_params(CalderaParams(Q_bits=4, L_bits=16, R_bits=8, rank=32, iters=3, lplr_iters=2,
                                     update_order=["LR", "Q"], sigma_reg=1e-3, rand_svd=True), True, seed=7)
    assert (cp.q_bits, cp.l_bits, cp.r_bits, cp.rank, cp.iters, cp.lplr_iters) == (4, 16, 8, 32, 3, 2)
    assert cp.n_order == 2 and list(cp.order)[:2] == [1, 0] and cp.rand_svd == 1 and cp.seed == 7
    assert abs(cp.sigma_reg - 1e-3) < 1e-9 and cp.q_block == 0
    with pytest.raises(ValueError):
        make_c_params(CalderaParams(update_order=["X"]), True)
