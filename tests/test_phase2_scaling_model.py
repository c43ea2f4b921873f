"""Exact-arithmetic model of the packed phase-2 kernel's Cp scaling (csrc/kernels_phase2.cuh, PK = true).

The reference computes (float)((double)pressure * 12.0 * 12.0 / (double)qbar) (cpp/exec/psp_process.cpp:2479-2496), i.e.
RN32(RN64(144 p / q)).  The kernel evaluates K = 144 / q = Kh + Kl, h = RN32(p Kh), e = p Kh - h, c = RN32(p Kl + e),
out = RN32(h + c) and redoes a group with the IEEE division when |c| lies within 16 bit-pattern units of half an ulp of
h.  Here both are evaluated in exact rational arithmetic (no GPU): whenever the kernel's test does not fire, the two
results must be the same float."""
from fractions import Fraction

import numpy as np


def rn32(x: Fraction) -> np.float32:
    """Round an exact rational to the nearest float32, ties to even (normal range only)."""
    if x == 0:
        return np.float32(0.0)
    s = -1 if x < 0 else 1
    a = abs(x)
    e = a.numerator.bit_length() - a.denominator.bit_length()       # 2^(e-1) <= a < 2^(e+1)
    if Fraction(2) ** e > a:
        e -= 1
    sc = a / Fraction(2) ** (e - 23)                                  # in [2^23, 2^24)
    m, r = divmod(sc.numerator, sc.denominator)
    twice = 2 * r
    if twice > sc.denominator or (twice == sc.denominator and (m & 1)):
        m += 1
    return np.float32(s * float(m) * 2.0 ** (e - 23))


def kernel_scaling(p: np.float32, q: np.float32):
    K = 144.0 / float(q)                                               # double, as the kernel computes it
    Kh = np.float32(K)
    Kl = np.float32(K - float(Kh))
    fp, fKh, fKl = Fraction(float(p)), Fraction(float(Kh)), Fraction(float(Kl))
    h = rn32(fp * fKh)
    e = fp * fKh - Fraction(float(h))                                  # exact in float (one FMA)
    assert Fraction(float(rn32(e))) == e
    c = rn32(fp * fKl + e)
    out = rn32(Fraction(float(h)) + Fraction(float(c)))
    cb = int(np.float32(c).view(np.uint32)) & 0x7fffffff
    hb = int(np.float32(h).view(np.uint32)) & 0x7f800000
    near = (cb - hb + 0x0C000010) & 0xffffffff
    return out, near <= 32


def reference_scaling(p: np.float32, q: np.float32) -> np.float32:
    x = float(p) * 12.0 * 12.0          # exact in double
    return np.float32(x / float(q))      # RN64 then RN32


def test_float_float_scaling_equals_reference_when_not_flagged():
    rng = np.random.default_rng(11)
    n, flagged = 60000, 0
    qs = rng.uniform(20.0, 3000.0, n).astype(np.float32)
    ps = (rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 1, n)).astype(np.float32)
    for p, q in zip(ps, qs):
        out, flag = kernel_scaling(p, q)
        flagged += flag
        if not flag:
            assert out.view(np.uint32) == reference_scaling(p, q).view(np.uint32), (p, q)
    assert flagged < n // 1000           # the exact path must stay rare


def test_values_next_to_a_rounding_midpoint_are_flagged_or_equal():
    """Search, for a few q, the pressures whose scaled value comes closest to the midpoint of two floats."""
    rng = np.random.default_rng(12)
    for q in rng.uniform(50.0, 2000.0, 6).astype(np.float32):
        K = Fraction(144) / Fraction(float(q))
        best = []
        base = np.float32(rng.uniform(0.001, 0.5))
        bits = int(base.view(np.uint32))
        for d in range(40000):
            p = np.uint32(bits + d).view(np.float32)
            v = Fraction(float(p)) * K
            r = rn32(v)
            ulp = Fraction(2.0 ** (np.frexp(r)[1] - 24))
            dist = abs(abs(v - Fraction(float(r))) - ulp / 2) / ulp       # 0 = exactly on a midpoint
            best.append((dist, p))
        best.sort(key=lambda t: t[0])
        for dist, p in best[:60]:
            out, flag = kernel_scaling(p, q)
            if not flag:
                assert out.view(np.uint32) == reference_scaling(p, q).view(np.uint32), (p, q, float(dist))
