"""Host-side checks (no GPU) for the SURVEY 8(f) additions: argument validation, the "no CPU fallback" rule,
and the dense Hadamard construction the GPU test compares against."""
import numpy as np
import pytest
import torch

from ee274_convexcaldera_llm_quantization_b200.hadamard import hadamard_transform, next_power_of_two
from ee274_convexcaldera_llm_quantization_b200.hessian import HessianAccumulator
from ee274_convexcaldera_llm_quantization_b200.linear import packed_linear


def test_next_power_of_two_matches_reference_definition():
    # main.py:75-77: 1 << (n - 1).bit_length()
    assert [next_power_of_two(n) for n in (1, 2, 3, 64, 65, 4096, 11008, 28672)] == [1, 2, 4, 64, 128, 4096, 16384, 32768]


def test_sylvester_construction_is_scipy_hadamard():
    scipy_linalg = pytest.importorskip("scipy.linalg")
    from test_gpu_hadamard import _sylvester      # tests/ is on sys.path (rootdir conftest)
    for n in (1, 2, 8, 64):
        assert np.array_equal(_sylvester(n), scipy_linalg.hadamard(n) / np.sqrt(n))     # main.py:79-83


def test_no_cpu_fallback_for_the_next_rows():
    x = torch.zeros(4, 64)
    with pytest.raises(ValueError):
        hadamard_transform(x)
    with pytest.raises(ValueError):
        HessianAccumulator(64, device="cpu")
    with pytest.raises(ValueError):
        packed_linear(x, torch.zeros(16, dtype=torch.uint8), torch.ones(1), 2, 1)
    with pytest.raises(ValueError):
        hadamard_transform(x, inverse=True)          # original_shape is required, checked before the device
