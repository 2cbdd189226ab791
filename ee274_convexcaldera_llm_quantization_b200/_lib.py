"""ctypes binding of libcaldera_b200.so (the C ABI declared in include/caldera_b200.h).

There is no CPU fallback: importing this module without the built library, or calling
into it without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# CB_LIBRARY selects another build of the same sources (scripts/probe_*.py use libcaldera_b200_measure.so, built with
# -DCB_MEASURE by `python -m ee274_convexcaldera_llm_quantization_b200.build --measure`)
LIB_PATH = os.path.join(_PKG, os.environ.get("CB_LIBRARY", "libcaldera_b200.so"))

CB_OK = 0
CB_ERR_ARG, CB_ERR_BITS, CB_ERR_BLOCK, CB_ERR_WORKSPACE, CB_ERR_UNSUPPORTED = -1, -2, -3, -4, -5
CB_H_IDENTITY, CB_H_DIAG, CB_H_DENSE = 0, 1, 2


class CalderaLibraryMissing(ImportError):
    pass


class cb_caldera_params(C.Structure):
    _fields_ = [
        ("compute_q", C.c_int32), ("compute_lr", C.c_int32),
        ("q_bits", C.c_int32), ("l_bits", C.c_int32), ("r_bits", C.c_int32),
        ("rank", C.c_int32), ("iters", C.c_int32), ("lplr_iters", C.c_int32),
        ("aware", C.c_int32), ("n_order", C.c_int32), ("order", C.c_int32 * 8),
        ("rand_svd", C.c_int32), ("sigma_reg", C.c_float), ("scale_w", C.c_int32),
        ("global_scale_in", C.c_float), ("q_block", C.c_int64),
        ("sketch_width", C.c_int32), ("power_iters", C.c_int32), ("power_iters_warm", C.c_int32), ("warm_start", C.c_int32),
        ("use_tensor_cores", C.c_int32), ("exec_mode", C.c_int32), ("seed", C.c_uint64),
    ]


class cb_caldera_out(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in (
        "Q", "L", "R", "Q_idxs", "Q_scale", "Q_packed", "L_idxs", "R_idxs", "L_scale", "R_scale",
        "L_packed", "R_packed", "W_scaled", "errors", "seed_dev", "scalars")]


_SIGNATURES = {
    "cb_version": (C.c_int, []),
    "cb_status_string": (C.c_char_p, [C.c_int]),
    "cb_kernel_launch_count": (C.c_int64, []),
    "cb_note_launches": (None, [C.c_int64]),
    "cb_quantize_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int64,
                                  C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cb_dequantize_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64,
                                    C.c_void_p, C.c_void_p]),
    "cb_packed_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "cb_pack_codes": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "cb_unpack_codes": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "cb_hessian_probe": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cb_weighted_error": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "cb_weighted_error_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "cb_lowrank_init": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                  C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_size_t, C.c_void_p]),
    "cb_lowrank_init_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "cb_lplr_iter_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int]),
    "cb_lplr_iter": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                               C.c_void_p]),
    "cb_cholesky_inverse_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cb_jacobi_eigh_from_chol_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p]),
    "cb_min_eig_shift_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "cb_min_eig_shift_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "cb_sgemm_strided": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                   C.c_int, C.c_void_p]),
    "cb_gemm_bf16_tn_workspace_bytes": (C.c_size_t, []),
    "cb_gemm_bf16_tn": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "cb_gemm_bf16_tn_bf16out": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "cb_gemm_bf16_tn_batched": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_int64,
                                          C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                          C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cb_quantize_nf_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_float, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "cb_dequantize_nf_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p]),
    "cb_hadamard_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "cb_hadamard_transform_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                            C.c_void_p, C.c_size_t, C.c_void_p]),
    "cb_hessian_accumulate_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "cb_hessian_accumulate_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_size_t, C.c_void_p]),
    "cb_packed_linear_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "cb_packed_linear_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p]),
    "cb_convert_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "cb_sum_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "cb_convex_prox_iters": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_float, C.c_int64, C.c_int64, C.c_int, C.c_uint64,
                                       C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.c_void_p]),
    "cb_convex_prox_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "cb_convex_prox_iters_dense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float,
                                             C.c_float, C.c_float, C.c_float, C.c_int64, C.c_int64, C.c_int, C.c_uint64,
                                             C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_size_t, C.c_void_p]),
    "cb_quantize_bbint_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cb_bbint_outliers_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "cb_dequantize_bbint_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "cb_convex_dense_prepare": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_size_t, C.c_void_p]),
    "cb_quantize_residual_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cb_scale_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cb_caldera_layer_workspace_bytes": (C.c_size_t, [C.POINTER(cb_caldera_params), C.c_int64, C.c_int64, C.c_int]),
    "cb_caldera_batch_supported": (C.c_int, [C.POINTER(cb_caldera_params), C.c_int64, C.c_int64, C.c_int]),
    "cb_caldera_batch": (C.c_int, [C.POINTER(cb_caldera_params), C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_int, C.POINTER(cb_caldera_out), C.c_void_p, C.c_size_t, C.c_void_p]),
    "cb_caldera_layer": (C.c_int, [C.POINTER(cb_caldera_params), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                   C.c_int, C.POINTER(cb_caldera_out), C.c_void_p, C.c_size_t, C.c_void_p]),
}

# measurement aids: exported by libcaldera_b200_measure.so only (include/caldera_b200.h, CB_MEASURE section)
_MEASURE_SIGNATURES = {
    "cb_set_gemm_staged_epilogue": (None, [C.c_int]),
    "cb_set_gemm_kblocks": (None, [C.c_int]),
    "cb_set_gemm_timing": (None, [C.c_void_p]),
    "cb_set_chol_timing": (None, [C.c_void_p]),
    "cb_probe_mma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cb_probe_err_pass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Returns the loaded library (cached).  Raises CalderaLibraryMissing if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CalderaLibraryMissing(
            f"{LIB_PATH} is missing: build it with `python -m ee274_convexcaldera_llm_quantization_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in _MEASURE_SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib


_MODES = {"latency": 0, "throughput": 1}
_mode = "latency"


def set_execution_mode(mode: str) -> None:
    """Default execution mode that `make_c_params` writes into cb_caldera_params.exec_mode: "latency" (a single layer
    finishes as early as possible) or "throughput" (many layers in flight on different streams; see
    include/caldera_b200.h).  Host-side default only: the library itself keeps no mode -- it is part of every call."""
    global _mode
    if mode not in _MODES:
        raise ValueError(f"execution mode must be one of {sorted(_MODES)}, got {mode!r}")
    _mode = mode


def execution_mode() -> str:
    return _mode


def execution_mode_code(mode=None) -> int:
    return _MODES[mode if mode is not None else _mode]


def exported_symbols():
    return sorted(_SIGNATURES)


def status_string(status: int) -> str:
    return load().cb_status_string(int(status)).decode()


class CalderaRuntimeError(RuntimeError):
    def __init__(self, status: int, where: str = ""):
        self.status = status
        super().__init__(f"{where}: libcaldera_b200 status {status} ({status_string(status)})")


def check(status: int, where: str = "") -> None:
    """Maps C status codes to the exception classes the reference raises."""
    if status == CB_OK:
        return
    if status == CB_ERR_BITS:
        raise AssertionError("Bit-width not supported!")
    if status == CB_ERR_BLOCK:
        raise ValueError(f"{where}: {status_string(status)}")
    if status == CB_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{where}: {status_string(status)}")
    if status == CB_ERR_ARG:
        raise ValueError(f"{where}: {status_string(status)}")
    raise CalderaRuntimeError(status, where)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what="tensor"):
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} must live on a CUDA device: this is the B200 path of the CALDERA hot loop and has no CPU "
            "fallback (use the reference implementation for CPU runs)")
