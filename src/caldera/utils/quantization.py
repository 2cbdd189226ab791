"""Reference path RCR/caldera/utils/quantization.py -> ee274_convexcaldera_llm_quantization_b200.quantization."""
from ee274_convexcaldera_llm_quantization_b200.quantization import (  # noqa: F401
    _BITWIDTHS, _QUANTIZER_METHODS, AbstractQuantizer, LowMemoryQuantizer, QuantizerFactory,
    pack_codes, unpack_codes)
