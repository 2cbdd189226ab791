"""Drop-in for RCR/caldera/utils/quantization.py (uniform method) on libcaldera_b200.

Same names, constructor arguments, return conventions and exceptions as the reference's
`LowMemoryQuantizer` / `QuantizerFactory` (quantization.py:18-37, 244-319).  The uniform
method runs in the sm_100a kernels of csrc/quant.cu; the NormalFloat codebooks (nf4 / nf2,
quantization.py:39-94) and the bitsandbytes-style asymmetric methods with an outlier side table
(bbint4 / bbint2, quantization.py:107-243; csrc/bbint.cu) are available through the same class
(SURVEY.md section 8f rank 4).  The CSV log of outlier counts the reference appends to in the
working directory (:126-137) is not written; `caldera()` itself accepts the uniform method only.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from typing import List

import torch

from . import _lib

_BITWIDTHS = [2, 4, 8, 16]
_QUANTIZER_METHODS = ["uniform", "nf4", "nf2", "bbint4", "bbint2"]
# quantization.py:44-50 and :57-60
_NF4_LEVELS = [-1.334, -1.0, -0.784, -0.617, -0.476, -0.347, -0.226, -0.112, 0.0, 0.112, 0.226, 0.347, 0.476, 0.617,
               0.784, 1.0]
_NF2_LEVELS = [-0.8165, -0.3333, 0.3333, 0.8165]


class AbstractQuantizer(ABC):
    @abstractmethod
    def quantize_block(self, weight): ...

    @abstractmethod
    def dequantize_block(self, weight_quant, weight_params, weight_shape): ...


class LowMemoryQuantizer(AbstractQuantizer):
    """quantization.py:18.  `quantize_block` returns (codes, scales, shape) with codes
    int8 (bits<=8) / int16 of shape (numel/block, block) and scales fp32 (numel/block, 1)."""

    def __init__(self, num_bits: int = 2, method: str = "uniform", block_size: int = 64):
        self.num_bits = num_bits
        assert self.num_bits in _BITWIDTHS, "Bit-width not supported!"
        self.method = method.lower()
        if self.method not in _QUANTIZER_METHODS:
            raise NotImplementedError(f"Quantization method '{self.method}' not supported yet.")
        self.block_size = block_size
        # same constructor-time checks as quantization.py:36-37, 42-43, 53-54
        if self.method == "nf4" and self.num_bits != 4:
            raise ValueError("NF4 quantization supports only 4 bits.")
        if self.method == "nf2" and self.num_bits != 2:
            raise ValueError("NF2 quantization supports only 2 bits.")
        if self.method == "bbint4" and self.num_bits != 4:
            raise ValueError("bbint4 quantization supports only 4 bits.")

        if self.method in ("nf4", "nf2"):
            # quantization.py:39-66: codebook and the thresholds midway between neighbouring levels (fp32)
            levels = _NF4_LEVELS if self.method == "nf4" else _NF2_LEVELS
            self.levels = torch.tensor(levels, dtype=torch.float32)
            self.thresholds = (self.levels[:-1] + self.levels[1:]) / 2

    # -- helpers -----------------------------------------------------------------
    def _check_method(self, what):
        if self.method not in _QUANTIZER_METHODS:
            raise NotImplementedError(f"{what} method '{self.method}' not implemented.")

    def _quantize_bbint(self, weight: torch.Tensor, epsilon: float):
        """quantization.py:107-154 / :175-221: (packed uint8 (numel/block, block*bits/8),
        (block_min, scales, outlier_values, outlier_indices)).  bbint2 packs 2-bit codes whatever num_bits says,
        like the reference (only bbint4 checks its bit width, :36-37)."""
        lib = _lib.load()
        bits = 4 if self.method == "bbint4" else 2
        bs = int(self.block_size)
        if bs < 2 or bs % (8 // bits) != 0:
            raise ValueError(f"{self.method}: block size {bs} must be a multiple of {8 // bits} (whole packed bytes)")
        w = (weight if weight.dtype == torch.float32 else weight.float()).contiguous()
        total = w.numel()
        nblk = total // bs
        dev = w.device
        packed = torch.empty((nblk, bs * bits // 8), dtype=torch.uint8, device=dev)
        block_min = torch.empty((nblk, 1), dtype=torch.float32, device=dev)
        scales = torch.empty((nblk, 1), dtype=torch.float32, device=dev)
        counts = torch.empty(nblk, dtype=torch.int32, device=dev)
        offsets = torch.empty(nblk + 1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.cb_quantize_bbint_f32(_lib.ptr(w), total, bs, bits, float(epsilon), _lib.ptr(packed),
                                                 _lib.ptr(block_min), _lib.ptr(scales), _lib.ptr(counts), _lib.ptr(offsets),
                                                 _lib.stream_ptr()), "quantize_block")
            n_out = int(offsets[-1].item())                 # sizes the side table (the one host synchronisation)
            values = torch.empty(n_out, dtype=torch.float32, device=dev)
            indices = torch.empty((n_out, 2), dtype=torch.int64, device=dev)
            if n_out > 0:
                _lib.check(lib.cb_bbint_outliers_f32(_lib.ptr(w), total, bs, float(epsilon), _lib.ptr(counts), _lib.ptr(offsets),
                                                     _lib.ptr(values), _lib.ptr(indices), _lib.stream_ptr()), "quantize_block")
        return packed, (block_min, scales, values, indices)

    def _dequantize_bbint(self, weight_packed: torch.Tensor, params, weight_shape):
        """quantization.py:156-173 / :223-243."""
        lib = _lib.load()
        bits = 4 if self.method == "bbint4" else 2
        block_min, scales, values, indices = params
        q = weight_packed.contiguous()
        numel = q.numel() * (8 // bits)
        nblk = scales.numel()
        out = torch.empty(numel, dtype=torch.float32, device=q.device)
        values = values.to(q.device, torch.float32).contiguous()
        indices = indices.to(q.device, torch.int64).contiguous()
        with torch.cuda.device(q.device):
            _lib.check(lib.cb_dequantize_bbint_f32(_lib.ptr(q), _lib.ptr(block_min.contiguous().float()),
                                                   _lib.ptr(scales.contiguous().float()), numel, numel // nblk, bits,
                                                   _lib.ptr(values) if values.numel() else None,
                                                   _lib.ptr(indices) if values.numel() else None, values.numel(),
                                                   _lib.ptr(out), _lib.stream_ptr()), "dequantize_block")
        return out.reshape(tuple(weight_shape))

    def _quantize_nf(self, weight: torch.Tensor, epsilon: float):
        """quantization.py:270-279 with _quantize_nf (:68-88): uint8 level indices, fp32 block scales."""
        lib = _lib.load()
        w = (weight if weight.dtype == torch.float32 else weight.float()).contiguous()    # the reference flattens
        total = w.numel()
        nblk = total // self.block_size
        idx = torch.empty((nblk, self.block_size), dtype=torch.uint8, device=w.device)
        scales = torch.empty((nblk, 1), dtype=torch.float32, device=w.device)
        thr = (C.c_float * len(self.thresholds))(*[float(t) for t in self.thresholds])
        with torch.cuda.device(w.device):
            st = lib.cb_quantize_nf_f32(_lib.ptr(w), total, self.block_size, thr, len(self.thresholds), float(epsilon),
                                        _lib.ptr(idx), _lib.ptr(scales), _lib.stream_ptr())
        _lib.check(st, "quantize_block")
        return idx, scales, weight.shape

    def _dequantize_nf(self, weight_quant: torch.Tensor, weight_params, weight_shape):
        """quantization.py:296-298 with _dequantize_nf (:90-94)."""
        lib = _lib.load()
        numel = 1
        for s_ in weight_shape:
            numel *= int(s_)
        q = weight_quant.contiguous()
        scales = weight_params.contiguous().float()
        out = torch.empty(numel, dtype=torch.float32, device=q.device)
        lv = (C.c_float * len(self.levels))(*[float(t) for t in self.levels])
        with torch.cuda.device(q.device):
            st = lib.cb_dequantize_nf_f32(_lib.ptr(q), _lib.ptr(scales), numel, numel // scales.numel(), lv,
                                          len(self.levels), _lib.ptr(out), _lib.stream_ptr())
        _lib.check(st, "dequantize_block")
        return out.reshape(tuple(weight_shape))

    def quantize_block(self, weight: torch.Tensor, epsilon: float = 1e-8, return_packed: bool = False):
        """quantization.py:244-288.  With return_packed=True a fourth value, the packed
        uint8 code stream (MSB-first, offset binary), is appended.  return_packed="only" skips the int8 / int16 codes
        altogether and returns (packed, scales, shape) -- the quantise + pack operation of the wire format, one pass at
        4 + bits/8 + 4/block bytes per element; dequantize_block accepts the packed stream."""
        if len(weight.shape) != 2:
            raise ValueError(f"Support only for 2D matrix, but your input has {len(weight.shape)} dimensions.")
        total = weight.shape[0] * weight.shape[1]
        if total % self.block_size != 0:
            raise ValueError(
                f"Weight with shape {weight.shape[0]} x {weight.shape[1]} "
                f"is not divisible by block size {self.block_size}")
        self._check_method("Quantization")
        _lib.require_cuda(weight, "weight")
        if self.method in ("nf4", "nf2"):
            if return_packed:
                raise NotImplementedError("return_packed is available for method='uniform' only")
            return self._quantize_nf(weight, epsilon)
        if self.method in ("bbint4", "bbint2"):
            if return_packed:
                raise NotImplementedError("return_packed is available for method='uniform' only (bbint codes are always packed)")
            weight_quant, weight_params = self._quantize_bbint(weight, epsilon)
            return weight_quant, weight_params, weight.shape
        lib = _lib.load()
        w = weight if weight.dtype == torch.float32 else weight.float()
        nblk = total // self.block_size
        cdtype = torch.int8 if self.num_bits <= 8 else torch.int16
        pack_only = isinstance(return_packed, str)
        if pack_only and return_packed != "only":
            raise ValueError('return_packed must be False, True or "only"')
        codes = None if pack_only else torch.empty((nblk, self.block_size), dtype=cdtype, device=w.device)
        scales = torch.empty((nblk, 1), dtype=torch.float32, device=w.device)
        packed = None
        if return_packed:
            packed = torch.empty(lib.cb_packed_bytes(total, self.num_bits), dtype=torch.uint8, device=w.device)
        with torch.cuda.device(w.device):
            st = lib.cb_quantize_f32(_lib.ptr(w), w.shape[0], w.shape[1], w.stride(0), w.stride(1),
                                     self.num_bits, self.block_size, float(epsilon), _lib.ptr(codes),
                                     _lib.ptr(packed), _lib.ptr(scales), None, _lib.stream_ptr())
        _lib.check(st, "quantize_block")
        if pack_only:
            return packed, scales, weight.shape
        if return_packed:
            return codes, scales, weight.shape, packed
        return codes, scales, weight.shape

    def dequantize_block(self, weight_quant: torch.Tensor, weight_params, weight_shape: List[int]):
        """quantization.py:290-307.  `weight_quant` may be the int8/int16 codes or the packed
        uint8 stream produced with return_packed=True."""
        self._check_method("Dequantization")
        _lib.require_cuda(weight_quant, "weight_quant")
        if self.method in ("nf4", "nf2"):
            return self._dequantize_nf(weight_quant, weight_params, weight_shape)
        if self.method in ("bbint4", "bbint2"):
            return self._dequantize_bbint(weight_quant, weight_params, weight_shape)
        lib = _lib.load()
        numel = 1
        for s in weight_shape:
            numel *= int(s)
        scales = weight_params.contiguous().float()
        out = torch.empty(numel, dtype=torch.float32, device=weight_quant.device)
        q = weight_quant.contiguous()
        is_packed = q.dtype == torch.uint8
        with torch.cuda.device(q.device):
            st = lib.cb_dequantize_f32(None if is_packed else _lib.ptr(q), _lib.ptr(q) if is_packed else None,
                                       _lib.ptr(scales), numel, self.num_bits, numel // scales.numel(),
                                       _lib.ptr(out), _lib.stream_ptr())
        _lib.check(st, "dequantize_block")
        return out.reshape(tuple(weight_shape))


class QuantizerFactory:
    """quantization.py:310-319."""

    def __init__(self, method="uniform", block_size=64):
        self.method = method
        self.block_size = block_size

    def get_quantizer(self, num_bits, device="cpu"):
        return LowMemoryQuantizer(num_bits=num_bits, method=self.method, block_size=self.block_size)

    def __str__(self):
        return f"QuantizerFactory(method={self.method}, block_size={self.block_size})"


def pack_codes(codes: torch.Tensor, num_bits: int) -> torch.Tensor:
    """int8/int16 codes -> packed uint8 stream (layout in include/caldera_b200.h)."""
    _lib.require_cuda(codes, "codes")
    lib = _lib.load()
    c = codes.contiguous()
    out = torch.empty(lib.cb_packed_bytes(c.numel(), num_bits), dtype=torch.uint8, device=c.device)
    with torch.cuda.device(c.device):
        _lib.check(lib.cb_pack_codes(_lib.ptr(c), c.numel(), num_bits, _lib.ptr(out), _lib.stream_ptr()), "pack_codes")
    return out


def unpack_codes(packed: torch.Tensor, num_bits: int, numel: int) -> torch.Tensor:
    _lib.require_cuda(packed, "packed")
    lib = _lib.load()
    out = torch.empty(numel, dtype=torch.int8 if num_bits <= 8 else torch.int16, device=packed.device)
    with torch.cuda.device(packed.device):
        _lib.check(lib.cb_unpack_codes(_lib.ptr(packed.contiguous()), numel, num_bits, _lib.ptr(out),
                                       _lib.stream_ptr()), "unpack_codes")
    return out
