"""CPU tests: pin the oracle against the reference's own fixtures / golden vectors
(SURVEY 8c).  No GPU needed."""
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN


def test_unpack12_matches_reference_python_reader_on_mraw_fixture(orc):
    g = np.load(os.path.join(GOLDEN, "mraw_golden.npz"))
    got = orc.unpack_12bit(g["packed_head"])
    assert np.array_equal(got, g["pixels_head"])
    assert got.max() <= 4095
    assert np.array_equal(orc.pack_12bit(got), g["packed_head"])   # round trip


def test_unpack_pack_roundtrip_10_and_12(orc):
    rng = np.random.default_rng(0)
    p12 = rng.integers(0, 4096, 4096).astype(np.uint16)
    assert np.array_equal(orc.unpack_12bit(orc.pack_12bit(p12)), p12)
    p10 = rng.integers(0, 1024, 4096).astype(np.uint16)
    assert np.array_equal(orc.unpack_10bit(orc.pack_10bit(p10)), p10)


def test_warp_affine_matches_cv2_golden(orc):
    g = np.load(os.path.join(GOLDEN, "warp_golden.npz"))
    for f in range(g["src"].shape[0]):
        assert np.array_equal(orc.warp_affine(g["src"][f], g["m6"][f], 1), g["linear"][f])
        assert np.array_equal(orc.warp_affine(g["src"][f], g["m6"][f], 0), g["nearest"][f])
    g = np.load(os.path.join(GOLDEN, "warp_f32_golden.npz"))
    for f in range(g["src"].shape[0]):
        assert np.array_equal(orc.warp_affine(g["src"][f], g["m6"][f], 1), g["linear"][f])


def test_warp_affine_matches_live_cv2(orc):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for t in range(6):
        h, w = 70 + t, 90 - t
        img = rng.integers(0, 4096, (h, w)).astype(np.uint16)
        M = np.eye(2, 3, dtype=np.float32)
        M[:, :2] += rng.normal(0, 1e-2, (2, 2)).astype(np.float32)
        M[:, 2] = rng.normal(0, 3, 2).astype(np.float32)
        for interp, flag in ((1, cv2.INTER_LINEAR), (0, cv2.INTER_NEAREST)):
            ref = cv2.warpAffine(img, M, (w, h), flags=flag | cv2.WARP_INVERSE_MAP)
            assert np.array_equal(orc.warp_affine(img, M, interp), ref)


def test_hot_pixel_rules(orc):
    img = np.full((8, 8), 1000, np.uint16)
    img[3, 3] = 4095
    out, n = orc.fix_hot_pixels(img)
    assert n == 1 and out[3, 3] == 1000
    img[3, 3] = 1500          # drop of 500 <= 512 but not hot (below 4064)
    out, n = orc.fix_hot_pixels(img)
    assert n == 0 and np.array_equal(out, img)
    img[:] = 1000
    img.reshape(-1)[[1, 9, 17, 25, 33, 41]] = 4095     # six hot pixels: untouched
    out, n = orc.fix_hot_pixels(img)
    assert n == -1 and np.array_equal(out, img)
    img[:] = 3800
    img[0, 0] = 4064          # corner: 2 neighbours, upper median; drop 264 -> kept
    out, n = orc.fix_hot_pixels(img)
    assert n == 1 and out[0, 0] == 4064


def test_detrend_reference_unit_test_recipe(orc):
    """cpp/test/test_filtering.cpp:19-85"""
    F, npts = 25, 13
    x = (np.arange(F, dtype=np.float32) / np.float32(F))
    for p in range(npts):
        y = np.zeros(F, np.float32)
        for c in range(7):
            y += (np.power(x.astype(np.float64), c) * np.float32(2.5 / (c + 1) + p / (c + 1))).astype(np.float32)
        fit, _ = orc.transpoly_fit(y, 6)
        assert np.abs(fit - y).max() < 1e-4


def test_colpiv_qr_solves_well_conditioned_system(orc):
    rng = np.random.default_rng(1)
    A = rng.normal(size=(40, 6)).astype(np.float32)
    x = rng.normal(size=6).astype(np.float32)
    got = orc.colpiv_qr_solve(A, A @ x)
    assert np.abs(got - x).max() < 1e-4


def test_patch_polyfit_near_origin_and_pivot_truncation(orc):
    """Near the image origin the absolute-coordinate float cubic system (patches.ipp:183-203)
    has cond ~3e7: Eigen's ColPivHouseholderQR::solve keeps only nonzeroPivots() pivots
    (those above max-column-norm * eps), so the reference's own patched values are a
    rank-9 fit here -- within 1 % of a bilinear field, not exact.  Pixels outside the
    interior list are untouched."""
    h, w = 40, 40
    y, x = np.mgrid[0:h, 0:w]
    img = (1000 + 3.0 * x + 2.0 * y + 0.05 * x * y).astype(np.float32)
    import upsp_b200
    b, i = upsp_b200.synth.make_patches(h, w, n_targets=1, seed=0)
    b, i = b[:1], i[:1]
    out = orc.patch_apply(img, orc.Patches(b, i))
    ix, iy = i[0]
    assert np.abs(out[iy, ix] - img[iy, ix]).max() < 0.01 * img.max()
    untouched = np.ones((h, w), bool)
    untouched[iy, ix] = False
    assert np.array_equal(out[untouched], img[untouched])
    # a cluster with fewer than 10 boundary pixels is skipped (patches.ipp:112-115)
    few = [(b[0][0][:9], b[0][1][:9])]
    assert np.array_equal(orc.patch_apply(img, orc.Patches(few, i)), img)


def test_transpose_and_apportion(orc):
    """transpose_block-style exactness (cpp/test/test_general_utils.cpp:32-100) for the
    rank-simulated global_transpose, ragged sizes."""
    rng = np.random.default_rng(2)
    for (F, N, R) in ((7, 5, 1), (103, 57, 3), (250, 301, 4), (5, 9, 8)):
        a = rng.normal(size=(F, N)).astype(np.float32)
        fs, fe = orc.apportion(F, R)
        assert fe.sum() == F and np.all(np.diff(fs) == fe[:-1]) and fe.max() - fe.min() <= 1
        parts = orc.global_transpose([a[fs[r]:fs[r] + fe[r]] for r in range(R)], N, F)
        assert np.array_equal(np.concatenate(parts, 0), a.T)


def test_overlap_remap_equals_reference_loop(orc):
    import upsp_b200
    n = 500
    ov = upsp_b200.synth.make_overlap(n, 60, seed=9)
    src = orc.overlap_remap(n, ov)
    sol = np.random.default_rng(3).normal(size=n).astype(np.float32)
    ref = sol.copy()
    for curr in sorted(ov):                       # P3DModel.ipp:144-157 verbatim semantics
        for alt in ov[curr]:
            if curr < alt:
                ref[alt] = ref[curr]
    assert np.array_equal(sol[src], ref)


def test_phase2_matches_float64_model(orc):
    """float-QR oracle vs float64 least squares: the reference's own float noise is << 1e-5
    of the operand scale."""
    rng = np.random.default_rng(4)
    n, F = 20, 400
    it = (1800 + rng.normal(0, 8, (n, F))).astype(np.float32)
    avg = it.mean(1).astype(np.float32)
    cov = np.ones(n, np.float32)
    cov[3] = 0
    import upsp_b200
    cal, qbar, ps, steady, temp = upsp_b200.synth.tunnel_conditions(n)
    p32, rms, av, gain = orc.phase2(it, avg, cov, steady, temp, cal, qbar, ps)
    p64, *_ = orc.phase2(it, avg, cov, steady, temp, cal, qbar, ps, exact_fit=True)
    assert np.isnan(rms[3]) and np.all(p32[3] == 0)
    K = np.abs(gain) * 144.0 / qbar
    ok = cov != 0
    assert (np.abs(p32 - p64)[ok].max(1) / K[ok]).max() < 1e-5
    assert np.allclose(gain[ok], [orc.get_gain(cal, temp[i], qbar * steady[i] + ps) for i in np.nonzero(ok)[0]])


def test_ecc_restatement_matches_cv2_golden(orc):
    """oracle/ecc.py (numpy restatement of cv::findTransformECC) against cv2.findTransformECC
    results stored by tests/golden/make_golden.py: same correlation to 1e-6, translation to
    2e-3 px, linear part to 2e-5 (OpenCV inverts the 6x6 Hessian in float LU; the restatement in
    double: that is the whole difference)."""
    from oracle import ecc
    g = np.load(os.path.join(GOLDEN, "ecc_golden.npz"))
    fr = g["frames"]
    ref32 = fr[0].astype(np.float32)
    for f in range(1, fr.shape[0]):
        M, rho, it = ecc.find_transform_ecc(ref32, fr[f].astype(np.float32))
        Mc = g["m6"][f - 1].reshape(2, 3)
        assert abs(rho - g["rho"][f - 1]) < 1e-6
        assert np.abs(M[:, 2] - Mc[:, 2]).max() < 2e-3
        assert np.abs(M[:, :2] - Mc[:, :2]).max() < 2e-5
        # registration recovers the synthetic jitter (inverse map: opposite sign)
        assert np.abs(M[:, 2] + g["shifts"][f]).max() < 0.06
        assert 2 <= it <= 8


def test_ecc_blur_and_gradient_match_live_cv2(orc):
    cv2 = pytest.importorskip("cv2")
    from oracle import ecc
    rng = np.random.default_rng(0)
    # integer-valued f32 (a converted 12-bit frame, what register_pixel feeds ECC): every product
    # with k/16 and every partial sum is exact in float, so the summation order cannot matter
    img = rng.integers(0, 4096, (37, 53)).astype(np.float32)
    b = cv2.GaussianBlur(img, (5, 5), 0)
    assert np.array_equal(b, ecc.gaussian_blur5(img))
    gx = cv2.filter2D(b, -1, np.array([[-0.5, 0, 0.5]], np.float32))
    gy = cv2.filter2D(b, -1, np.array([[-0.5], [0], [0.5]], np.float32))
    ogx, ogy = ecc.gradients(b)
    assert np.array_equal(gx, ogx) and np.array_equal(gy, ogy)


def test_spatial_filter_matches_live_cv2(orc):
    """a5: GaussianBlur / blur on CV_16U and CV_32F (12-bit data) equal cv2 bit for bit."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    img = rng.integers(0, 4096, (41, 59)).astype(np.uint16)
    f = img.astype(np.float32)
    for ks in (3, 5, 7):
        assert np.array_equal(orc.spatial_filter(img, 1, ks), cv2.GaussianBlur(img, (ks, ks), 0))
        assert np.array_equal(orc.spatial_filter(img, 2, ks), cv2.blur(img, (ks, ks)))
        assert np.array_equal(orc.spatial_filter(f, 1, ks), cv2.GaussianBlur(f, (ks, ks), 0))
        assert np.array_equal(orc.spatial_filter(f, 2, ks), cv2.blur(f, (ks, ks)))
    # non-integer f32 (next to patched pixels): within one ulp of cv2 (its SIMD path may fuse mul-add)
    g = f + rng.random(f.shape).astype(np.float32)
    for ks in (3, 5, 7):
        a, b = orc.spatial_filter(g, 1, ks), cv2.GaussianBlur(g, (ks, ks), 0)
        assert np.abs(a - b).max() <= 2.5e-7 * np.abs(b).max()
    # CV_16U Gaussians beyond the three small kernels: OpenCV's error-diffused fixed-point taps, any odd size
    full = rng.integers(0, 65536, (33, 47)).astype(np.uint16)
    for ks in (1, 9, 11, 15, 21, 31):
        for im in (img, full):
            assert np.array_equal(orc.spatial_filter(im, 1, ks), cv2.GaussianBlur(im, (ks, ks), 0)), ks
    assert np.array_equal(orc.spatial_filter(f, 1, 1), f) and np.array_equal(orc.spatial_filter(f, 2, 1), f)
    with pytest.raises(ValueError):
        orc.spatial_filter(f, 1, 9)            # CV_32F beyond 7: depends on OpenCV's SIMD summation order, not restated


def test_fixed_point_gaussian_taps_match_cv2():
    """csrc/gauss_fixed.inc (the taps k_filter_u16 uses for deck filter = gaussian, any odd size 3..31): the integer
    filter  (sum_j q_j sum_i q_i p + 2^31) >> 32  with these taps equals cv2.GaussianBlur on CV_16U bit for bit."""
    cv2 = pytest.importorskip("cv2")
    import re
    from conftest import ROOT
    text = open(os.path.join(ROOT, "upsp-processing_b200", "csrc", "gauss_fixed.inc")).read()
    rows = [list(map(int, m.split(","))) for m in re.findall(r"^\s*\{([0-9, ]+)\},", text, re.M)]
    assert len(rows) == 15
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 4096, (61, 83)).astype(np.uint16), rng.integers(0, 65536, (40, 37)).astype(np.uint16)]
    for k, half in zip(range(3, 33, 2), rows):
        assert len(half) == k // 2 + 1
        q = np.array(half + half[-2::-1], np.int64)
        assert q.sum() == 65536
        for img in imgs:
            r = k // 2
            p = np.pad(img.astype(np.int64), r, mode="reflect")          # BORDER_REFLECT_101
            H, W = img.shape
            rows_ = sum(q[i] * p[:, i:i + W] for i in range(k))
            S = sum(q[j] * rows_[j:j + H] for j in range(k))
            got = np.clip((S + (1 << 31)) >> 32, 0, 65535).astype(np.uint16)
            assert np.array_equal(got, cv2.GaussianBlur(img, (k, k), 0)), k
