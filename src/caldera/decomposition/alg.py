"""Reference path RCR/caldera/decomposition/alg.py -> ee274_convexcaldera_llm_quantization_b200.alg."""
from ee274_convexcaldera_llm_quantization_b200.params import *  # noqa: F401,F403
from ee274_convexcaldera_llm_quantization_b200.quantization import *  # noqa: F401,F403
from ee274_convexcaldera_llm_quantization_b200.alg import caldera, activation_aware_error  # noqa: F401
