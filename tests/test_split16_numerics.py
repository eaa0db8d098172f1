"""CPU check of the arithmetic behind MAGNET_SRC_SPLIT16 (the tensor-core kernel's operand format):
x*s = hi + lo with two fp16 planes and a power-of-two scale from the largest finite |x|; the kernel accumulates
hi*hi + hi*lo + lo*hi in fp32 (products of fp16 numbers are exact there).  Restated in numpy: how far is that from the
exact dot product?  (The GPU side of the same claim: scripts/debug_mma.py compares the accumulator rows with an fp64
all-pairs product; the parity tests bound the end result.)"""
import numpy as np
import pytest


def split16(x):
    """The split of csrc/cost_mma.cu (split16_shift / split16_repack_kernel), in numpy."""
    x = np.asarray(x, dtype=np.float32)
    fin = np.isfinite(x)
    amax = float(np.abs(x[fin]).max()) if fin.any() else 0.0
    shift = 0
    if amax > 0:
        e = int(np.floor(np.log2(amax)))                  # amax in [2^e, 2^(e+1))
        shift = max(-100, min(100, 14 - e))
    s = np.float32(2.0) ** shift
    v = (x * s).astype(np.float32)                        # exact: power of two
    hi = v.astype(np.float16)
    with np.errstate(invalid="ignore"):                  # inf - inf = NaN, on purpose
        lo = (v - hi.astype(np.float32)).astype(np.float16)   # v - hi is exact in fp32
    return hi, lo, float(s)


@pytest.mark.parametrize("scale", [1e-3, 1.0, 37.0, 1e3])
def test_split_represents_fp32_to_22_bits(scale):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(1 << 16) * scale).astype(np.float32)
    hi, lo, s = split16(x)
    assert np.isfinite(hi.astype(np.float32)).all() and np.abs(hi.astype(np.float32)).max() < 65504
    assert 2.0 ** 14 <= np.abs(x).max() * s < 2.0 ** 15
    rec = (hi.astype(np.float64) + lo.astype(np.float64)) / s
    big = np.abs(x) >= np.abs(x).max() * 2.0 ** -18      # below that the lo plane runs into fp16 subnormals
    rel = np.abs(rec - x.astype(np.float64))[big] / np.abs(x.astype(np.float64))[big]
    assert rel.max() <= 2.0 ** -21
    assert np.abs(rec - x)[~big].max() <= np.abs(x).max() * 2.0 ** -38 if (~big).any() else True


def test_three_products_match_the_exact_dot_product():
    """64-channel dot products as the kernel forms them: exact products of the split factors, three of the four cross
    terms, summed — against the fp64 dot product of the fp32 inputs.  Error bound 2^-20 of sum |a||b| (the dropped
    lo*lo term is 2^-22; fp32 accumulation of the tensor core adds its own ~2^-23 per term, not modelled here)."""
    rng = np.random.default_rng(2)
    a = rng.standard_normal((512, 64)).astype(np.float32) * 3
    b = rng.standard_normal((700, 64)).astype(np.float32) * 0.2
    ah, al, sa = split16(a)
    bh, bl, sb = split16(b)
    f = lambda t: t.astype(np.float64)
    g = (f(ah) @ f(bh).T + f(ah) @ f(bl).T + f(al) @ f(bh).T) / (sa * sb)
    exact = f(a) @ f(b).T
    bound = np.abs(f(a)) @ np.abs(f(b)).T
    assert (np.abs(g - exact) <= 2.0 ** -20 * bound).all()
    # and relative to the largest dot product (what the parity bar measures): far inside 1e-4
    assert np.abs(g - exact).max() <= 1e-6 * np.abs(exact).max()


def test_non_finite_elements_do_not_set_the_scale():
    x = np.array([1.0, -3.0, np.inf, np.nan, 0.5], dtype=np.float32)
    hi, lo, s = split16(x)
    assert s == 2.0 ** 13                                  # from |-3|: 3 * 2^13 in [2^14, 2^15)
    assert np.isinf(hi[2]) and np.isnan(lo[2]) and np.isnan(hi[3])          # inf poisons its own products (inf - inf)
    assert float(hi[0]) + float(lo[0]) == 2.0 ** 13 and float(hi[1]) + float(lo[1]) == -3 * 2.0 ** 13
    z_hi, z_lo, zs = split16(np.zeros(4, dtype=np.float32))
    assert zs == 1.0 and not z_hi.any() and not z_lo.any()
