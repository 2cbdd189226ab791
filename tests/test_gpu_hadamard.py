"""Fast Walsh-Hadamard pre-rotation against the dense fp64 construction of the reference (main.py:79-133)."""
import numpy as np
import pytest
import torch

from ee274_convexcaldera_llm_quantization_b200.hadamard import hadamard_transform, next_power_of_two

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sylvester(n):
    """scipy.linalg.hadamard(n) / sqrt(n) (main.py:79-83)."""
    H = np.array([[1.0]])
    while H.shape[0] < n:
        H = np.block([[H, H], [H, -H]])
    return H / np.sqrt(n)


@pytest.mark.parametrize("rows,cols", [(64, 64), (100, 300), (1, 17), (896, 1152), (2048, 4096)])
def test_hadamard_matches_dense_reference(rows, cols):
    g = torch.Generator().manual_seed(rows + cols)
    W = torch.randn(rows, cols, generator=g)
    out, shape = hadamard_transform(W.to(DEV))
    pr, pc = next_power_of_two(rows), next_power_of_two(cols)
    assert shape == (rows, cols) and tuple(out.shape) == (pr, pc)
    padded = np.zeros((pr, pc))
    padded[:rows, :cols] = W.numpy()
    ref = _sylvester(pr) @ padded @ _sylvester(pc)            # main.py:118-121
    assert float(np.abs(out.cpu().numpy() - ref).max()) < 2e-5 * max(1.0, float(np.abs(ref).max()))
    back = hadamard_transform(out, inverse=True, original_shape=shape)
    assert float((back.cpu() - W).abs().max()) < 2e-5
    np.testing.assert_allclose(float(out.double().norm()), float(W.double().norm()), rtol=1e-5)   # orthogonal: energy preserved
