"""Reference path RCR/caldera/utils/dataclasses.py -> ee274_convexcaldera_llm_quantization_b200.params."""
from ee274_convexcaldera_llm_quantization_b200.params import *  # noqa: F401,F403
from ee274_convexcaldera_llm_quantization_b200.params import CalderaParams, CalderaDecomposition, QuantInfo  # noqa: F401
from ee274_convexcaldera_llm_quantization_b200.quantization import (  # noqa: F401
    QuantizerFactory, AbstractQuantizer, LowMemoryQuantizer)
