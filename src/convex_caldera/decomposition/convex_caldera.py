"""Reference path RCR/convex_caldera/decomposition/convex_caldera.py ->
ee274_convexcaldera_llm_quantization_b200.convex_caldera."""
from ee274_convexcaldera_llm_quantization_b200.convex_caldera import (  # noqa: F401
    ConvexCalderaParams, ConvexCalderaDecomposition, convex_caldera, round_bit_allocations)
from ee274_convexcaldera_llm_quantization_b200.params import CalderaDecomposition  # noqa: F401
