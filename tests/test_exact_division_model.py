"""CPU model of the division the quantiser kernels use for bit widths above 2 (csrc/common.cuh: make_scale_recip /
div_by_scale_exact): y = RN(1 / s), q0 = RN(x y), two residual corrections q <- fma(fma(-q, s, x), y, q).  The kernels
rely on the result being the correctly rounded quotient RN(x / s) -- what the reference's `x / s` computes
(quantization.py:266) -- whenever ScaleRecip::exact holds.  Here every fused multiply-add is evaluated in exact rational
arithmetic and rounded once to fp32 (ties to even), and the result is compared with numpy's IEEE division."""
from fractions import Fraction

import numpy as np

f32 = np.float32


def _round_f32(v: Fraction) -> np.float32:
    """Nearest fp32 to an exact rational, ties to even (no double rounding: neighbours are compared exactly)."""
    if v == 0:
        return f32(0.0)
    c = f32(float(v))
    cands = {c, np.nextafter(c, f32(-np.inf)), np.nextafter(c, f32(np.inf))}
    best, best_err = None, None
    for k in cands:
        if not np.isfinite(k):
            continue
        err = abs(Fraction(float(k)) - v)
        even = (int(np.float32(k).view(np.uint32)) & 1) == 0
        if best is None or err < best_err or (err == best_err and even):
            best, best_err = k, err
    return f32(best)


def _fma(a, b, c) -> np.float32:
    return _round_f32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def _div_model(x, s):
    y = f32(1.0) / s                                     # __frcp_rn
    q0 = _round_f32(Fraction(float(x)) * Fraction(float(y)))
    q1 = _fma(_fma(-q0, s, x), y, q0)
    return _fma(_fma(-q1, s, x), y, q1)


def _exact_ok(s) -> bool:                                 # ScaleRecip::exact
    return 1e-18 < s < 1e18 and (int(np.float32(s).view(np.uint32)) & 0x7FFFFF) != 0x7FFFFF


def test_reciprocal_multiply_with_two_corrections_is_the_ieee_quotient():
    rng = np.random.default_rng(17)
    mant = rng.integers(0, 1 << 23, size=300, dtype=np.uint32)
    expo = rng.integers(127 - 55, 127 + 55, size=300, dtype=np.uint32)
    scales = list(((expo << 23) | mant).view(np.float32)) + [f32(1.0), f32(3.0), f32(1e-8), f32(0.02), f32(7.0), f32(127.0)]
    checked = 0
    for s in scales:
        s = f32(s)
        if not _exact_ok(s):
            continue
        xs = list((rng.uniform(-1.0, 1.0, size=12).astype(np.float32) * s).astype(np.float32))
        # quotients next to the 4-bit decision boundaries (k + 1/2) / 7 and at the ends of the range
        for k in (0, 1, 3, 6):
            b = f32((k + 0.5) / 7.0) * s
            xs += [b, np.nextafter(b, f32(np.inf)), np.nextafter(b, f32(-np.inf)), -b]
        xs += [s, -s, f32(0.0), np.nextafter(s, f32(0.0))]
        for x in xs:
            x = f32(x)
            if abs(x) > s:
                continue
            want = f32(x) / s                               # IEEE, correctly rounded
            got = _div_model(x, s)
            assert got == want or (got == 0 and want == 0), (x, s, got, want)
            checked += 1
    assert checked > 5000
