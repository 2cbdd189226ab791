"""CPU check of the identity behind the division-free 2-bit paths (csrc/common.cuh: ternary_code; csrc/quant.cu: the
pack-only accumulation; csrc/stages.cu: quant_form_y_body<TERN>): for fp32 x with |x| <= s and a scale inside the
validity range,

    rint(fl(fl(x / s) * 1)) == (x > s/2) - (x < -s/2)                      (quantization.py:266, levels = 1)

and the packed symbol code + 1 == (x > s/2) + !(x < -s/2).  Exhaustive around the decision boundaries (every fp32
neighbour of +-s/2 within a few ulps) for scales of many magnitudes and significand patterns, plus random interiors."""
import numpy as np

from oracle import caldera_oracle as orc

f32 = np.float32


def _scales():
    rng = np.random.default_rng(5)
    s = [f32(1.0), f32(0.5), f32(3.0), f32(1e-8), f32(1e-29), f32(9e29), f32(0.02), f32(6.693187), f32(1.9999999),
         np.nextafter(f32(2.0), f32(0.0)), np.nextafter(f32(1.0), f32(2.0)), f32(2.0 ** -99), f32(2.0 ** 99)]
    mant = rng.integers(0, 1 << 23, size=400, dtype=np.uint32)
    expo = rng.integers(127 - 98, 127 + 98, size=400, dtype=np.uint32)
    s += list(((expo << 23) | mant).view(np.float32))
    s += list((((expo[:50]) << 23) | np.uint32(0x7FFFFF)).view(np.float32))      # all-ones significands
    return np.array(s, dtype=np.float32)


def _candidates(s):
    """x values for one scale: +-s/2 and their neighbours, +-s, zeros, tiny values, random interiors."""
    hs = f32(0.5) * s
    assert f32(2.0) * hs == s                                   # s / 2 is exact
    xs = [f32(0.0), f32(-0.0), s, -s, f32(1e-38), f32(-1e-38)]
    for centre in (hs, -hs):
        lo = hi = centre
        xs.append(centre)
        for _ in range(6):
            lo = np.nextafter(lo, f32(-np.inf))
            hi = np.nextafter(hi, f32(np.inf))
            xs += [lo, hi]
    rng = np.random.default_rng(int(np.float32(s).view(np.uint32)) & 0xFFFF)
    xs += list((rng.uniform(-1.0, 1.0, size=64).astype(np.float32) * s).astype(np.float32))
    x = np.array(xs, dtype=np.float32)
    return x[np.abs(x) <= s]


def test_compare_only_code_equals_the_reference_arithmetic():
    for s in _scales():
        assert 1e-30 < s < 1e30                                  # ternary_ok
        x = _candidates(s)
        hs = f32(0.5) * s
        ref = np.rint((x / s).astype(np.float32) * f32(1.0)).astype(np.int32)      # IEEE divide, multiply, half-even
        code = (x > hs).astype(np.int32) - (x < -hs).astype(np.int32)
        np.testing.assert_array_equal(code, ref, err_msg=f"scale {s!r}")
        sym = (x > hs).astype(np.int32) + (~(x < -hs)).astype(np.int32)
        np.testing.assert_array_equal(sym, ref + 1)
        # dequantised value of the ternary grid: (code / 1) * s is the exact product
        np.testing.assert_array_equal((code.astype(np.float32) / f32(1.0)) * s, code.astype(np.float32) * s)


def test_accumulated_symbols_equal_the_packed_stream():
    """The pack-only kernel adds unit * [(x > s/2) + !(x < -s/2)] in fp32 (FSET + FFMA): the sum stays below 2^16, so it
    is exact, and equals the oracle's MSB-first packing of the reference codes (quantization.py:217-220)."""
    rng = np.random.default_rng(11)
    x = (rng.standard_normal((64, 256)) * np.exp(rng.standard_normal((64, 256)))).astype(np.float32)
    x[3, :64] = 0.0
    x[4, :8] = np.array([1.0, 0.5, -0.5, 0.25, -0.25, 0.75, -1.0, 0.5000001], dtype=np.float32)
    codes, scales, _ = orc.quantize_uniform(x, 2, 64)
    want = orc.pack_codes(codes, 2).reshape(-1)
    xb = x.reshape(-1, 64)
    hs = (f32(0.5) * scales).astype(np.float32)
    got = []
    for blk in range(xb.shape[0]):
        for lane in range(8):                                    # a lane owns 8 consecutive elements = two packed bytes
            acc = f32(0.0)
            for e in range(8):
                unit = f32(1 << (6 - 2 * e if e < 4 else 14 - 2 * (e - 4)))
                v = xb[blk, lane * 8 + e]
                acc = f32(acc + unit * f32(v > hs[blk, 0]))
                acc = f32(acc + unit * f32(not (v < -hs[blk, 0])))
            h = int(acc)
            got += [h & 255, h >> 8]
    np.testing.assert_array_equal(np.array(got, dtype=np.uint8), want)
