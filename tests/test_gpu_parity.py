"""GPU parity tests: the CUDA library (through its C ABI) against the CPU oracle on the same
seeded inputs.  Integer / index / byte work and the whole of phase 1 must be BIT-EXACT;
delta-Cp time histories are held to 1e-5 relative (fp32) as defined in chain.cp_errors."""
import numpy as np
import pytest

from chain import Case, cp_errors, degenerate_nodes, monomial_mass, run_gpu, run_oracle, same_bits

pytestmark = pytest.mark.gpu

CP_TOL = 1e-5    # BASELINE.json north_star: "within 1e-5 relative (fp32) on per-node Cp time histories"
EXACT_TOL = 1e-6  # GPU vs the float64 least-squares model of the same chain (10x tighter)


# ------------------------------------------------------------------ stand-alone operators
def test_unpack12_matches_oracle(up, orc, gpu):
    rng = np.random.default_rng(0)
    for npix in (1024 * 1024, 1280 * 800, 6, 8 * 37 + 2):
        pix = rng.integers(0, 4096, npix).astype(np.uint16)
        packed = orc.pack_12bit(pix)
        got = up.op_unpack(packed, up.PIX_PACKED12, npix)
        assert np.array_equal(got, orc.unpack_12bit(packed))
        assert np.array_equal(got, pix)


def test_unpack10_with_and_without_lut(up, orc, gpu):
    rng = np.random.default_rng(1)
    npix = 64 * 48
    pix = rng.integers(0, 1024, npix).astype(np.uint16)
    packed = orc.pack_10bit(pix)
    lut = (np.arange(1024, dtype=np.uint32) * 4 + 1).astype(np.uint16)
    assert np.array_equal(up.op_unpack(packed, up.PIX_PACKED10, npix), orc.unpack_10bit(packed))
    assert np.array_equal(up.op_unpack(packed, up.PIX_PACKED10, npix, lut), orc.unpack_10bit(packed, lut))
    assert np.array_equal(up.op_unpack(packed, up.PIX_PACKED10, npix), pix)


def _hot_cases(rng, h, w):
    base = rng.integers(100, 3000, (h, w)).astype(np.uint16)
    cases = []
    for pts in ([], [(5, 5)], [(0, 0)], [(h - 1, w - 1)], [(0, 7), (h - 1, 3), (4, 0), (9, w - 1)],
                [(3, 3), (3, 4)],                      # adjacent hot pixels: the second sees the first's fix
                [(1, 1), (2, 2), (3, 3), (4, 4), (5, 5)],           # exactly max_hot
                [(1, 1), (2, 2), (3, 3), (4, 4), (5, 5), (6, 6)],   # too many: untouched
                ):
        img = base.copy()
        for (r, c) in pts:
            img[r, c] = 4095
        cases.append(img)
    img = base.copy()
    img[10, 10] = 4064          # at the threshold
    img[9, 10] = img[11, 10] = img[10, 9] = img[10, 11] = 3800   # drop < 512: kept
    cases.append(img)
    img = base.copy()
    img[20, 20] = 4063          # just below: not hot
    cases.append(img)
    return np.stack(cases)


def test_fix_hot_pixels_matches_oracle(up, orc, gpu):
    rng = np.random.default_rng(2)
    frames = _hot_cases(rng, 32, 40)
    got, n_hot = up.op_fix_hot_pixels(frames)
    for f in range(frames.shape[0]):
        ref, n = orc.fix_hot_pixels(frames[f])
        assert np.array_equal(got[f], ref), f"frame {f}"
        assert n_hot[f] == n


@pytest.mark.parametrize("interp", [0, 1])
def test_warp_affine_matches_oracle_and_golden(up, orc, gpu, interp):
    import os
    from conftest import GOLDEN
    rng = np.random.default_rng(3)
    h, w = 97, 130
    nf = 12
    frames = rng.integers(0, 4096, (nf, h, w)).astype(np.uint16)
    m = np.zeros((nf, 2, 3), np.float32)
    m[:, 0, 0] = m[:, 1, 1] = 1
    scale = np.array([5e-4, 5e-2, 0.3])[np.arange(nf) % 3]
    m[:, :, :2] += (rng.normal(0, 1, (nf, 2, 2)) * scale[:, None, None]).astype(np.float32)
    m[:, :, 2] = (rng.normal(0, 1, (nf, 2)) * np.array([1, 5, 60])[np.arange(nf) % 3][:, None]).astype(np.float32)
    got = up.op_warp_affine(frames, m.reshape(nf, 6), interp)
    for f in range(nf):
        assert np.array_equal(got[f], orc.warp_affine(frames[f], m[f], interp)), f"frame {f}"
    # the committed cv2.warpAffine golden vectors (tests/golden/make_golden.py)
    g = np.load(os.path.join(GOLDEN, "warp_golden.npz"))
    got = up.op_warp_affine(g["src"], g["m6"], interp)
    assert np.array_equal(got, g["linear" if interp == 1 else "nearest"])


def test_project_frames_matches_oracle(up, orc, gpu):
    import upsp_b200
    rng = np.random.default_rng(4)
    h, w, n = 64, 80, 5000
    frames = rng.uniform(0, 4095, (5, h, w)).astype(np.float32)
    for csr in (upsp_b200.synth.make_projection(n, h, w, "random", weights=True),
                upsp_b200.synth.make_multi_nnz_projection(n, h, w, 9)):
        got = up.op_project_frames(*csr, frames)
        for f in range(frames.shape[0]):
            assert same_bits(got[f], orc.project_frame(*csr, frames[f]))


@pytest.mark.parametrize("shape", [(1, 1), (7, 5), (64, 64), (100, 257), (333, 64), (129, 1000)])
def test_transpose_exact(up, orc, gpu, shape):
    rng = np.random.default_rng(5)
    a = rng.normal(size=shape).astype(np.float32)
    assert np.array_equal(up.op_transpose(a), a.T)


def test_detrend_reference_unit_test_recipe(up, orc, gpu):
    """cpp/test/test_filtering.cpp:19-85: degree-6 data reproduced to 1e-4 abs (F=25, 13 pts)."""
    F, npts = 25, 13
    x = (np.arange(F, dtype=np.float32) / np.float32(F))
    y = np.zeros((npts, F), np.float32)
    for p in range(npts):
        for c in range(7):
            y[p] += (np.power(x.astype(np.float64), c) * np.float32(2.5 / (c + 1) + p / (c + 1))).astype(np.float32)
    fit = up.op_polyfit_detrend(y, 6)
    assert np.abs(fit - y).max() < 1e-4
    for p in range(npts):
        ofit, _ = orc.transpoly_fit(y[p], 6)
        assert np.abs(fit[p] - ofit).max() < 1e-4


@pytest.mark.parametrize("F", [25, 1000, 20000, 40000, 70001, 410001])
def test_detrend_long_series_vs_oracle(up, orc, gpu, F):
    """ratio-like series (1 + drift + noise): GPU fit vs the float-QR oracle fit and vs the
    float64 least-squares fit.  F=40000 / 70001 run as 2- / 4-CTA clusters (row split over
    distributed shared memory); F=410001 exceeds 8 x shared memory (two-HBM-pass path)."""
    rng = np.random.default_rng(6)
    t = np.arange(F) / F
    y = (1.0 + 0.02 * np.sin(2 * np.pi * t) + 0.01 * t + rng.normal(0, 4e-3, (3, F))).astype(np.float32)
    fit = up.op_polyfit_detrend(y, 6)
    A = np.stack([(np.arange(F, dtype=np.float32) / np.float32(F)).astype(np.float64) ** c for c in range(7)], 1)
    for p in range(3):
        exact = A @ np.linalg.lstsq(A, y[p].astype(np.float64), rcond=None)[0]
        assert np.abs(fit[p] - exact).max() < 2.5e-7          # ~2 ulp of 1.0
        if F <= 20000:
            ofit, _ = orc.transpoly_fit(y[p], 6)
            assert np.abs(fit[p] - ofit).max() < 1e-5


# ------------------------------------------------------------------ the whole chain
def _check_chain(case, ref, got, orc, strict=False):
    """Phase 1, transpose, gain: bit-exact.  delta-Cp: the detrend is the one stage the GPU does
    not compute in the reference's float operation order (DESIGN.md "tolerances"), so
      (a) vs the float64 least-squares model of the reference chain: <= 1e-6 of the operand
          scale K_n*max|r| on every node, plus 8 eps32 * sum|monomial coef| (the precision to
          which the reference's float design matrix defines the fit, chain.monomial_mass);
      (b) vs the float-QR oracle (the reference's arithmetic): <= 1e-5 plus that oracle's own
          distance from the float64 model on the node (its float rounding noise, which reaches
          ~1e-4 on short series with outliers -- no implementation that does not replay Eigen's
          exact float sequence can be closer than that);
      (c) strict=True (clean series, config 1): plain <= 1e-5 vs the float-QR oracle."""
    if got["intensity"] is not None:
        assert same_bits(got["intensity"], ref["intensity"]), "frame-major intensity not bit-exact"
    assert same_bits(got["avg"], ref["avg"]) and same_bits(got["rms"], ref["rms"])
    assert same_bits(got["coverage"], ref["coverage"])
    assert same_bits(got["itrans"], ref["itrans"]), "intensity_transpose not bit-exact"
    assert same_bits(got["gain"], ref["gain"])
    skipped = ref["coverage"] == 0
    assert np.all(got["ptrans"][skipped] == 0) and np.all(np.isnan(got["rms2"][skipped]))
    deg = degenerate_nodes(ref)          # a zero intensity sample: the whole row is NaN in both
    assert not np.isfinite(ref["ptrans"][deg]).any() and not np.isfinite(got["ptrans"][deg]).any()
    exact = run_oracle(orc, case, exact_fit=True)
    e_exact, _ = cp_errors(case, exact, got)
    cond = 8 * np.finfo(np.float32).eps * monomial_mass(orc, ref)
    assert np.all(e_exact <= EXACT_TOL + cond), f"delta-Cp vs float64 model: {(e_exact - cond).max():.3e} > {EXACT_TOL}"
    e_op, e_cp = cp_errors(case, ref, got)
    noise, _ = cp_errors(case, ref, exact)       # the float-QR oracle's own rounding noise
    assert np.all(e_op <= CP_TOL + noise + cond), f"delta-Cp vs float-QR oracle: {(e_op - noise - cond).max():.3e} > {CP_TOL}"
    if strict:
        assert e_op.max() <= CP_TOL, f"delta-Cp error {e_op.max():.3e} (relative to operands) > {CP_TOL}"
    v = ~skipped & ~deg
    # statistics of delta-Cp: same criterion, against the float64 model and the float-QR oracle
    K = np.abs(ref["gain"][v]).astype(np.float64) * 144.0 / float(case.qbar)
    for key in ("rms2", "avg2"):
        assert np.all(np.abs(got[key][v] - exact[key][v]) / K <= EXACT_TOL + cond)
        nz = np.abs(ref[key][v] - exact[key][v]) / K
        assert np.all(np.abs(got[key][v] - ref[key][v]) / K <= CP_TOL + nz + cond)
    return e_op.max(), e_cp.max()


MODES = pytest.mark.parametrize("keep_frame_major", [False, True], ids=["fused", "frame-major+transpose"])


@MODES
def test_chain_config1_plain(up, orc, gpu, keep_frame_major):
    """config 1: one camera, registration none, patcher none, u16 frames, 128 frames of a still
    camera (clean series: the strict 1e-5 criterion against the float-QR oracle applies)."""
    import upsp_b200
    case = Case(upsp_b200.synth, n_frames=128, n_nodes=5000, jitter=False)
    ref = run_oracle(orc, case)
    got = run_gpu(up, orc, case, keep_frame_major=keep_frame_major)
    _check_chain(case, ref, got, orc, strict=True)
    assert got["launches"] > 0


@MODES
def test_chain_packed12_small_batches_and_ring(up, orc, gpu, keep_frame_major):
    """12-bit packed input, batch of 5 frames, 16-slot input ring (streaming mode)."""
    import upsp_b200
    case = Case(upsp_b200.synth, n_frames=50, n_nodes=2000, fmt="p12", seed=3)
    ref = run_oracle(orc, case)
    got = run_gpu(up, orc, case, batch_frames=5, frame_capacity=16, keep_frame_major=keep_frame_major)
    _check_chain(case, ref, got, orc)


def test_ring_lap_longer_than_the_record_ring(up, orc, gpu):
    """One frame per push / process_frames call through a 20-slot input ring: a lap of the ring spans 20 calls,
    more than the 16 process records the library keeps, so the push that overwrites a slot must fall back to the
    most recent process event (the evicted record used to mean "no wait")."""
    import upsp_b200
    from chain import push_all, setup_ctx
    case = Case(upsp_b200.synth, n_frames=70, n_nodes=2000, fmt="p12", registration=True, seed=5)
    ref = run_oracle(orc, case)
    g, sl = setup_ctx(up, orc, case, batch_frames=4, frame_capacity=20)
    push_all(up, orc, g, case, sl, chunk=1)
    g.finish_phase1()
    g.transpose()
    it = g.read_intensity_transpose()
    g.reset_run()                                  # a second run over the same ring: records of the first are void
    push_all(up, orc, g, case, sl, chunk=1)
    g.finish_phase1()
    g.transpose()
    it2 = g.read_intensity_transpose()
    g.close()
    assert same_bits(it, ref["itrans"]) and same_bits(it2, ref["itrans"])


@MODES
def test_chain_registration_patches_overlap(up, orc, gpu, keep_frame_major):
    """warp (given matrices) + polynomial patcher (incl. dependent clusters and a clipped
    target) + P3D overlap remap; linear and nearest interpolation."""
    import upsp_b200
    for interp in (1, 0):
        case = Case(upsp_b200.synth, n_frames=40, n_nodes=6000, registration=True, interp=interp,
                    patches=True, overlap=True, overlap_pair=True, kind="random", seed=7)
        ref = run_oracle(orc, case)
        got = run_gpu(up, orc, case, alias=False, keep_frame_major=keep_frame_major)
        _check_chain(case, ref, got, orc)


@MODES
def test_chain_multi_camera_weights(up, orc, gpu, keep_frame_major):
    """3 cameras, weighted entries, nodes seen by 0..3 cameras."""
    import upsp_b200
    case = Case(upsp_b200.synth, n_cams=3, n_frames=70, n_nodes=4000, registration=True, patches=True, seed=11)
    ref = run_oracle(orc, case)
    got = run_gpu(up, orc, case, keep_frame_major=keep_frame_major)
    _check_chain(case, ref, got, orc)


def test_chain_general_csr(up, orc, gpu):
    """rows with up to 4 entries (cfg-5 variant) take the general CSR kernel (never fused)."""
    import upsp_b200
    case = Case(upsp_b200.synth, n_cams=2, n_frames=20, n_nodes=3000, multi_nnz=4, seed=13)
    ref = run_oracle(orc, case)
    got = run_gpu(up, orc, case, keep_frame_major=True)
    _check_chain(case, ref, got, orc)


@MODES
def test_chain_ragged_sizes(up, orc, gpu, keep_frame_major):
    """odd frame size (not a multiple of 8), F and N not multiples of the tile sizes."""
    import upsp_b200
    case = Case(upsp_b200.synth, n_frames=37, n_nodes=1003, height=51, width=67, registration=True,
                patches=True, seed=17)
    ref = run_oracle(orc, case)
    got = run_gpu(up, orc, case, keep_frame_major=keep_frame_major)
    _check_chain(case, ref, got, orc)


def test_frame_major_read_needs_keep_flag(up, orc, gpu):
    import upsp_b200
    case = Case(upsp_b200.synth, n_frames=8, n_nodes=300)
    from chain import push_all, setup_ctx
    g, sl = setup_ctx(up, orc, case)
    push_all(up, orc, g, case, sl)
    with pytest.raises(up.UpspGpuError):
        g.read_intensity()
    g.close()


@pytest.mark.parametrize("kind,ksize", [(1, 3), (1, 5), (1, 7), (2, 3), (2, 5), (1, 9), (1, 15), (1, 1)])
@pytest.mark.parametrize("patches", [False, True], ids=["u16-image", "f32-image(patched)"])
def test_chain_spatial_filter(up, orc, gpu, kind, ksize, patches):
    """deck @options filter = gaussian | box (psp_process.cpp:1802-1807) after registration and
    patching: CV_16U image without the patcher (OpenCV's integer paths; Gaussian of any odd size), CV_32F with it
    (Gaussian 3 / 5 / 7: larger float kernels depend on OpenCV's SIMD summation order and are refused)."""
    import upsp_b200
    case = Case(upsp_b200.synth, n_frames=24, n_nodes=2500, registration=True, patches=patches, seed=29,
                filter_kind=kind, filter_size=ksize)
    if patches and kind == 1 and ksize > 7:
        with pytest.raises(up.UpspGpuError, match="patched"):
            run_gpu(up, orc, case, keep_frame_major=True)
        return
    ref = run_oracle(orc, case)
    got = run_gpu(up, orc, case, keep_frame_major=True)
    _check_chain(case, ref, got, orc)
    with pytest.raises(up.UpspGpuError):
        g = up.PspGpu(1, 10, 4)
        g.set_filter(1, 33)         # gaussian sizes beyond 31 are not built
    with pytest.raises(up.UpspGpuError):
        g = up.PspGpu(1, 10, 4)
        g.set_filter(2, 4)          # even sizes: psp_process.cpp:1296


@pytest.mark.parametrize("env", [{"UPSP_FUSED_V1": "1"}, {"UPSP_PIPELINE": "0"}, {"UPSP_PIPELINE": "0", "UPSP_DECODE_P": "1"}],
                         ids=["fused-v1", "no-pipeline", "persistent-decode-serial"])
def test_kernel_variants_match(up, orc, gpu, env):
    """The first-generation fused kernel (UPSP_FUSED_V1=1), the single-stream schedule
    (UPSP_PIPELINE=0) and the persistent decoder outside the pipeline must produce the same bits as
    the default path does against the oracle (subprocess: the knobs are read once).  Frames span
    several batches so that both buffer sets of the two-stream pipeline are exercised, one frame
    carries hot pixels, and the 96x128 frames keep npix % 32 == 0 so the persistent decoder is
    eligible."""
    import subprocess, sys, textwrap
    from conftest import ROOT
    code = textwrap.dedent("""
        import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
        import upsp_b200 as up
        from oracle import oracle as orc
        from chain import Case, run_gpu, run_oracle, same_bits
        for seed, ncam in ((7, 1), (8, 2)):
            case = Case(up.synth, n_frames=40, n_nodes=6000, n_cams=ncam, registration=True, patches=True,
                        overlap=True, seed=seed, fmt="p12")
            ref = run_oracle(orc, case); got = run_gpu(up, orc, case, batch_frames=8)
            assert same_bits(got["itrans"], ref["itrans"]) and same_bits(got["avg"], ref["avg"])
            assert same_bits(got["rms"], ref["rms"])
        print("variant ok")
    """ % (ROOT, ROOT))
    import os
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stderr[-2000:]


def test_phase2_mid_length_rows(up, orc, gpu):
    """Rows whose dynamic shared memory alone is below 48 KB but static + dynamic is above it
    (F = 8192: 32 KB + 18.5 KB) need the opt-in attribute as well: regression test for the
    'invalid argument' launch failure."""
    import upsp_b200
    case = Case(upsp_b200.synth, n_frames=8192, n_nodes=40, registration=False, seed=31, height=32, width=32)
    ref = run_oracle(orc, case)
    got = run_gpu(up, orc, case)
    _check_chain(case, ref, got, orc)


@pytest.mark.parametrize("registration,F", [(False, 32768), (True, 32768), (True, 98304)],
                         ids=["float-rows-2cta", "16bit-rows-2cta", "16bit-rows-8cta"])
def test_phase2_clustered_rows(up, orc, gpu, registration, F):
    """Rows too long for one CTA's shared memory (multi-GPU weak scaling makes them N_gpus times longer): the symmetric
    phase-2 kernel runs as a cluster of 2 / 8 CTAs per row: moments over distributed shared memory behind ONE cluster
    barrier, partial statistics through global memory (k_phase2_cl_parts).  With registration the projection takes the
    TMA path and the rows are 16-bit integers (+ float side rows for the unseen nodes)."""
    import upsp_b200
    case = Case(upsp_b200.synth, n_frames=F, n_nodes=24, registration=registration, seed=41, height=32, width=32,
                fmt="p12" if registration else "u16")
    ref = run_oracle(orc, case)
    got = run_gpu(up, orc, case)
    _check_chain(case, ref, got, orc)


def test_streamed_column_block_reads(up, orc, gpu):
    """upsp_gpu_read_intensity_transpose_block_async: column blocks read while later frames are
    still being processed equal the final intensity_transpose."""
    import upsp_b200
    from chain import setup_ctx
    case = Case(upsp_b200.synth, n_frames=64, n_nodes=700, registration=True, seed=23)
    g, sl = setup_ctx(up, orc, case, batch_frames=8)
    blocks = []
    for o in range(0, 64, 16):
        g.push_frames(0, case.frames[0][o:o + 16], up.PIX_U16, o, 16)
        g.process_frames(o, 16)
        buf = np.full((case.N, 20), -1.0, np.float32)       # pitch 20 > 16 columns
        g.read_intensity_transpose_block_async(0, case.N, o, 16, buf.ctypes.data, 20)
        blocks.append(buf)
    g.wait_reads()
    g.finish_phase1()
    g.transpose()
    full = g.read_intensity_transpose()
    for i, buf in enumerate(blocks):
        assert same_bits(buf[:, :16], full[:, 16 * i:16 * i + 16])
        assert np.all(buf[:, 16:] == -1.0)
    with pytest.raises(up.UpspGpuError):
        g.read_intensity_transpose_block_async(0, case.N, 0, 65, blocks[0].ctypes.data, 80)
    g.close()


def test_error_behaviour(up, gpu):
    g = up.PspGpu(1, 10, 4)
    with pytest.raises(up.UpspGpuError):
        g.process_frames()            # no camera / projection yet
    g.set_camera(0, 8, 8)
    with pytest.raises(up.UpspGpuError):
        g.set_projection(0, np.arange(11, dtype=np.int32), np.full(10, 64, np.int32), np.ones(10, np.float32))
    with pytest.raises(up.UpspGpuError):
        g.transpose()
    with pytest.raises(up.UpspGpuError):
        g.phase2([0] * 6, 1.0, 1.0, np.zeros(10), np.zeros(10))
    g.close()


def _tma_case(synth, n_frames=150, seed=11):
    """Hot configuration in miniature (one camera, packed 12-bit frames, bilinear registration, patches,
    seam remap) with frames that leave the staged-box fast path: a rotation / a scale too large for the
    box (every tap through global memory), shifts that push most of the box outside the image (zero
    fill = BORDER_CONSTANT), a shift by several pixels (the box follows the frame's own origin)."""
    case = Case(synth, n_frames=n_frames, n_nodes=9000, height=96, width=128, registration=True,
                patches=True, overlap=True, seed=seed, fmt="p12")
    w = case.warp[0].reshape(-1, 2, 3)
    c, s = np.float32(np.cos(0.2)), np.float32(np.sin(0.2))
    w[5] = [[c, -s, 9.0], [s, c, -14.0]]
    w[17, :, 2] = (40.0, -30.0)
    w[18] = [[1.3, 0.0, -11.0], [0.0, 1.3, -9.0]]
    w[33, :, 2] = (-3.7, 2.2)
    w[34, :, 2] = (-130.0, 5.0)          # everything outside
    if n_frames > 70:
        w[66, :, 2] = (0.4, 97.0)
    return case


@pytest.mark.parametrize("mode", ["v4", "tma16", "tma12"])
@pytest.mark.parametrize("batch,capacity", [(0, 0), (8, 0), (12, 24)], ids=["one-batch", "batch8", "ring24"])
def test_tma_projection_modes(up, orc, gpu, monkeypatch, mode, batch, capacity):
    """k_project_tma (boxes staged by TMA from the decoded frames / from the packed frames) against the
    oracle, bit for bit, and the mode that actually ran.  150 frames in one batch = table stages of
    64 + 64 + 22 frames (a 2-frame tail group); batch 8 exercises both buffer sets; ring24 streams the
    frames through a 24-slot input ring (the packed-source boxes are cut from the ring itself)."""
    import upsp_b200
    from chain import push_all, setup_ctx
    monkeypatch.setenv("UPSP_PROJ", mode)
    case = _tma_case(upsp_b200.synth)
    ref = run_oracle(orc, case)
    g, sl = setup_ctx(up, orc, case, batch_frames=batch, frame_capacity=capacity)
    push_all(up, orc, g, case, sl, chunk=capacity if capacity else None)
    assert g.projection_mode() == {"v4": 0, "tma16": 1, "tma12": 2}[mode]
    g.finish_phase1()
    avg, rms, cov = g.read_phase1_stats()
    g.transpose()
    it = g.read_intensity_transpose()
    g.close()
    assert same_bits(it, ref["itrans"])
    assert same_bits(avg, ref["avg"]) and same_bits(rms, ref["rms"])


@pytest.mark.parametrize("stream", ["0", "1"], ids=["rows-in-smem", "streaming-pair"])
def test_chain_16bit_rows_and_phase2_paths(up, orc, gpu, monkeypatch, stream):
    """Unit projection values + a frame count divisible by 8: the node-major rows are stored as 16-bit integers (float
    side rows for the patched / unseen nodes) and phase 2 reads them either into shared memory (k_phase2_sym<IN16>) or
    through the streaming pair k_phase2_moments / k_phase2_apply (the multi-GPU path for long rows, forced here).  The
    whole chain against the oracle, every output through the ABI's readers (which widen the rows to float)."""
    import upsp_b200
    monkeypatch.setenv("UPSP_PHASE2_STREAM", stream)
    case = Case(upsp_b200.synth, n_frames=200, n_nodes=6000, height=96, width=128, registration=True, patches=True,
                seed=17, fmt="p12")
    ref = run_oracle(orc, case)
    got = run_gpu(up, orc, case, batch_frames=64)
    _check_chain(case, ref, got, orc)


def test_chain_batch_blocked_rows_single_rank(up, orc, gpu, monkeypatch):
    """UPSP_BLOCKED=1: the batch-blocked 16-bit row layout of the multi-rank exchange on one rank (chain against the
    oracle; frames pushed and processed in calls that are not multiples of the batch)."""
    import upsp_b200
    from chain import push_all, setup_ctx
    monkeypatch.setenv("UPSP_BLOCKED", "1")
    case = Case(upsp_b200.synth, n_frames=200, n_nodes=6000, height=96, width=128, registration=True, patches=True,
                seed=18, fmt="p12")
    ref = run_oracle(orc, case)
    g, sl = setup_ctx(up, orc, case, batch_frames=64)
    push_all(up, orc, g, case, sl, chunk=40)            # calls of 40 frames: batches are cut at the 64-frame block edges
    g.finish_phase1()
    got = dict(intensity=None)
    got["avg"], got["rms"], got["coverage"] = g.read_phase1_stats()
    g.transpose()
    got["itrans"] = g.read_intensity_transpose()
    g.phase2(case.cal, case.qbar, case.ps, case.steady, case.temp, case.degree)
    got["ptrans"] = g.read_pressure_transpose()
    got["rms2"], got["avg2"], got["gain"] = g.read_phase2_stats()
    g.close()
    _check_chain(case, ref, got, orc)


def test_tma_projection_weighted_values(up, orc, gpu, monkeypatch):
    """Projection values other than 1.0 (a weighted single camera) take the float-statistics variant."""
    import upsp_b200
    for mode in ("tma16", "tma12"):
        monkeypatch.setenv("UPSP_PROJ", mode)
        case = Case(upsp_b200.synth, n_frames=40, n_nodes=5000, height=96, width=128, registration=True,
                    weights=True, seed=12, fmt="p12")
        ref = run_oracle(orc, case)
        got = run_gpu(up, orc, case)
        assert same_bits(got["itrans"], ref["itrans"]) and same_bits(got["avg"], ref["avg"]) and same_bits(got["rms"], ref["rms"])
