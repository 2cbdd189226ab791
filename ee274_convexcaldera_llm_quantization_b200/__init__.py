"""B200-native CALDERA decomposition hot path.

Public surface mirrors the reference's modules (RCR = rank-constrained-regression-main/src):
  RCR/caldera/decomposition/alg.py        -> .alg            (caldera, activation_aware_error)
  RCR/caldera/utils/quantization.py       -> .quantization   (QuantizerFactory, LowMemoryQuantizer)
  RCR/caldera/utils/dataclasses.py        -> .params         (CalderaParams, CalderaDecomposition, QuantInfo)
The same objects are importable under the reference's own paths (`src.caldera...`) through the
`src/` shim package at the repository root.  All arithmetic runs in libcaldera_b200.so
(csrc/*.cu, C ABI in include/caldera_b200.h); there is no CPU fallback.
"""
import os as _os

# Layers are independent and are kept in flight on separate CUDA streams (scheduler.decompose_layers,
# bench.py).  The CUDA driver maps streams onto 8 hardware work queues unless told otherwise, which
# falsely serialises layers beyond 8; ask for the maximum before the CUDA context is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .params import CalderaParams, CalderaDecomposition, QuantInfo
from .quantization import QuantizerFactory, LowMemoryQuantizer, AbstractQuantizer, pack_codes, unpack_codes

from ._lib import set_execution_mode, execution_mode

__all__ = ["set_execution_mode", "execution_mode", "CalderaParams", "CalderaDecomposition", "QuantInfo", "QuantizerFactory", "LowMemoryQuantizer",
           "AbstractQuantizer", "pack_codes", "unpack_codes", "caldera"]


def caldera(*args, **kwargs):
    from .alg import caldera as _caldera
    return _caldera(*args, **kwargs)
