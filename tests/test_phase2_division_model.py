"""Model of the 16-bit-row division of the packed phase-2 kernel (csrc/kernels_phase2.cuh, div_u16_2): r = avg / I with I an
integer in [1, 65535] is evaluated as rc = MUFU.RCP(I), q = RN(a rc), rem = RN(a - I q), r = RN(q + rem rc) -- no Newton step
on the reciprocal.  The reference divides in IEEE single precision (cpp/exec/psp_process.cpp:2480, the node loop's
`sol_avg_final[i] / intensity_transpose_buf[idx][f]`); here the three-instruction sequence is emulated exactly (float64
holds every intermediate product exactly; the few candidates for a double rounding are re-checked in rational arithmetic) with
the reciprocal off by up to 3 ulp, which covers the hardware approximation's stated error of 1 ulp."""
from fractions import Fraction

import numpy as np


def _rn32(x: Fraction) -> np.float32:
    if x == 0:
        return np.float32(0.0)
    s = -1 if x < 0 else 1
    y = abs(x)
    e = y.numerator.bit_length() - y.denominator.bit_length()
    if Fraction(2) ** e > y:
        e -= 1
    sc = y / Fraction(2) ** (e - 23)
    m, r = divmod(sc.numerator, sc.denominator)
    if 2 * r > sc.denominator or (2 * r == sc.denominator and (m & 1)):
        m += 1
    return np.float32(s * float(m) * 2.0 ** (e - 23))


def _count_bad(a, b, rc):
    a64, b64, rc64 = a.astype(np.float64), b.astype(np.float64), rc.astype(np.float64)
    q = (a64 * rc64).astype(np.float32)                                  # 24 x 24-bit product: exact in float64
    rem = (a64 - b64 * q.astype(np.float64)).astype(np.float32)          # exact difference, one rounding
    t = rem.astype(np.float64) * rc64                                    # exact
    got = (q.astype(np.float64) + t).astype(np.float32)                  # may round twice: mismatches re-checked below
    ref = (a64 / b64).astype(np.float32)                                 # float64 quotient then float: safe for division
    bad = 0
    for i in np.nonzero(got.view(np.uint32) != ref.view(np.uint32))[0]:
        exact = Fraction(float(q[i])) + Fraction(float(t[i]))
        if _rn32(exact).view(np.uint32) != _rn32(Fraction(float(a[i])) / Fraction(float(b[i]))).view(np.uint32):
            bad += 1
    return bad


def test_three_instruction_division_is_correctly_rounded_for_16bit_denominators():
    rng = np.random.default_rng(5)
    n = 500_000
    for trial in range(4):
        b = rng.integers(1, 65536, n).astype(np.float32)
        if trial % 2 == 0:
            a = (rng.uniform(1, 2, n) * 2.0 ** rng.integers(-20, 20, n)).astype(np.float32)
        else:        # numerators that put the quotient next to the midpoint of two floats
            k = rng.integers(2 ** 23, 2 ** 24, n).astype(np.float64) + 0.5
            a = (k * b.astype(np.float64)).astype(np.float32)
        rc0 = (1.0 / b.astype(np.float64)).astype(np.float32)
        for d in (-3, -1, 0, 1, 3):
            rc = (rc0.view(np.int32) + d).view(np.float32)
            assert _count_bad(a, b, rc) == 0


def test_all_denominators_with_a_fixed_numerator():
    b = np.arange(1, 65536, dtype=np.float32)
    for a0 in (1.0, 3.0, 1234.567, 65535.0, 0.001953125, 4095.75):
        a = np.full(b.shape, a0, np.float32)
        rc0 = (1.0 / b.astype(np.float64)).astype(np.float32)
        for d in (-2, 0, 2):
            assert _count_bad(a, b, (rc0.view(np.int32) + d).view(np.float32)) == 0


def test_zero_half_test_of_the_packed_denominators_is_exact():
    """k_phase2_sym<IN16> takes the fast division for a group of eight 16-bit intensities unless one of them is zero: the
    unsigned minimum of the eight halves (VIMNMX.U16x2), then ((m - 0x00010001) & ~m & 0x80008000) != 0 on the two halves
    left.  A borrow out of a zero low half can only add a second positive, so the test is exact."""
    lo = np.arange(65536, dtype=np.uint64)
    rng = np.random.default_rng(0)
    his = np.unique(np.concatenate([[0, 1, 2, 0x7FFF, 0x8000, 0x8001, 0xFFFE, 0xFFFF], rng.integers(0, 65536, 200)])).astype(np.uint64)
    for h in his:
        m = (h << np.uint64(16)) | lo
        t = ((m - np.uint64(0x00010001)) & np.uint64(0xFFFFFFFF)) & (~m & np.uint64(0xFFFFFFFF)) & np.uint64(0x80008000)
        assert np.array_equal(t != 0, (lo == 0) | (h == 0))
