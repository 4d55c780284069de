"""The segment index of the fused kernels' tone-curve table (csrc/risp_fused.cuh: gtm_lookup) as plain arithmetic.

`FFMA.RZ(x, 4, 2^20)` has ulp 1/8 at 2^20, so the float's mantissa is floor(32 x) and bits 3..5 of the word hold
floor(4 x): `(bits & 0x38) | table` is the address of the 8-byte entry of the pixel's segment, x == 1 selects entry 4
(the pass-through pixel of tools_origin.py:435's half-open segments) and a NaN selects entry 7.  This test restates that
with numpy (round-toward-zero emulated exactly in float64) for every float32 within 64 ulps of a knot, every power of two,
and a million random inputs -- no GPU involved."""
import numpy as np


def table_entry(x32):
    x = x32.astype(np.float64)
    s = 1048576.0 + 4.0 * x                      # exact in float64 (4x has 24 significant bits, offset 2^20)
    rz = np.floor(s * 8.0) / 8.0                 # round toward zero onto the float32 grid at 2^20 (ulp 1/8), positive values
    bits = rz.astype(np.float32).view(np.uint32)
    assert np.array_equal(rz.astype(np.float32).astype(np.float64), rz)      # representable: the emulation is exact
    return (bits & np.uint32(0x38)) >> np.uint32(3)


def neighbourhood(v, n=64):
    out = [np.float32(v)]
    lo = hi = np.float32(v)
    for _ in range(n):
        lo = np.nextafter(lo, np.float32(-1))
        hi = np.nextafter(hi, np.float32(2))
        out += [lo, hi]
    a = np.array(out, dtype=np.float32)
    return a[(a >= 0) & (a <= 1)]


def test_segment_index_is_floor_4x():
    rng = np.random.default_rng(0)
    xs = np.concatenate([neighbourhood(k) for k in (0.0, 0.25, 0.5, 0.75, 1.0)] +
                        [np.float32(2.0) ** -np.arange(0, 127, dtype=np.float32)] +
                        [rng.random(1 << 20, dtype=np.float32), np.array([1e-8, 1e-38, 1e-45], dtype=np.float32)])
    want = np.minimum(np.floor(4.0 * xs.astype(np.float64)), 4).astype(np.uint32)
    assert np.array_equal(table_entry(xs), want)
    assert table_entry(np.array([1.0], dtype=np.float32))[0] == 4
    # knots belong to the UPPER segment (half-open [x_k, x_k+1))
    assert table_entry(np.array([0.25, 0.5, 0.75], dtype=np.float32)).tolist() == [1, 2, 3]
    assert table_entry(np.nextafter(np.array([0.25, 0.5, 0.75, 1.0], dtype=np.float32), np.float32(0))).tolist() == [0, 1, 2, 3]


def test_nan_selects_the_pass_through_entry():
    # fma(NaN, 4, 2^20) is the canonical quiet NaN 0x7fffffff on the GPU: bits 3..5 are all set -> entry 7 = (a, s) = (0, 1)
    assert (0x7fffffff & 0x38) >> 3 == 7
