"""The product's host-side readers held against the REFERENCE'S OWN CODE: oracle/_ref/ref_probe is a small main() linked
with the reference sources that compile without third-party libraries (cpp/lib/upsp_inputs.cpp, non_cv_upsp.cpp, plot3d.cpp,
cpp/utils/general_utils.cpp, file_writers.cpp, file_io.cpp; `make -C oracle ref`).  Both probes print what they parsed in
the same line format; the outputs must be equal, for well-formed, odd and broken inputs alike.  CPU only."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN
from test_run_inputs import DOC_DECK, _p3d_function


@pytest.fixture(scope="module")
def probes(up, orc):
    ref = orc.build_ref()
    if ref is None or not os.path.exists(os.path.join(ref, "ref_probe")):
        pytest.skip("oracle/_ref/ref_probe not built and the reference tree is not present on this machine")
    # the probe is linked with --unresolved-symbols=ignore-all (the unused drawing code of cv_extras.cpp): make sure none of the
    # entry points compiled from reference lines (oracle/Makefile, ref_*) was left unresolved by a stale object
    nm = subprocess.run(["nm", "-C", os.path.join(ref, "ref_probe")], capture_output=True, text=True).stdout
    assert not [l for l in nm.splitlines() if " U ref_" in l or " U upsp::intensity_histc" in l or " U upsp::normal" in l or " U upsp::area" in l]
    return up.build.build_inputs_probe(), os.path.join(ref, "ref_probe")


def both(probes, *args):
    out = []
    for exe in probes:
        r = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True)
        lines = [l for l in r.stdout.splitlines() if not l.startswith(("version ", "Warning", "check 1"))]
        out.append((r.returncode != 0, lines))
    return out


DECKS = {
    "documented": DOC_DECK,
    "launcher": "%Version 9\n\n@general\n\ttest = t11\n\trun = 4121\n\tsequence = 07\n\ttunnel = ames_unitary\n\tframerate = 10000\n@all\n"
                "\tgrid = /g/m.tri\n\tsds = /g/r.wtd\n\ttargets = /g/m.tgts\n\tnormals = \n@camera\n\tnumber = 1\n\tcine = /g/a.cine\n"
                "\tcalibration = /c/1.json\n@options\n\ttarget_patcher = none\n\tregistration = none\n\tfilter = none\n\tfilter_size = 1\n"
                "\toblique_angle = 70\n\tnumber_frames = -1\n@output\n\tdir = /o\n\tname = 412107",
    "unsorted-cameras-own-targets": "@all\n grid = a.grd\n targets = all.tgts\n calibration = all.json\n@camera\n number = 3\n filename = c3.mraw\n"
                                    "@camera\n number = 1\n filename = c1.mraw\n targets = one.tgts\n@camera\n number = 2\n cine = c2.cine\n"
                                    " calibration = two.json\n",
    "odd-spacing-and-equals": "  # comment\n@general\n   test=a b c\n run =   12   \n sequence= 3x\n key = v = w\n novalue =\n = noname\n"
                              "@options\n  oblique_angle = 65.25e0\n overlap=best_view\n pixel_interpolation = nearest\n filter = box\n"
                              " filter_size = 5\n registration = pixel\n target_patcher = polynomial\n number_frames = 17\n@output\n name = n\n",
    "vars-twice-and-inside": "@vars\n root = /data/run\n tag = 0042\n@all\n grid = $root/g_$tag.x\n sds = $root/$tag/$tag.wtd\n"
                             " active_comps = pre$tag.csv\n@camera\n number = 1\n filename = $root/cam$tag.mraw\n@output\n dir = $root/out\n",
    "vars-second-in-the-middle": "@vars\n root = /data/run\n tag = 0042\n@all\n grid = $root/g.x\n sds = $root/$tag.wtd\n@output\n dir = $root/o/$tag\n",
    "var-prefix-clash": "@vars\n d = /x\n dir = /y\n@all\n grid = $dir/g.tri\n sds = $d/s.wtd\n",
    "at-sign-in-value": "@general\n test = me@host\n run = 4\n@all\n grid = g.p3d\n",
    "no-trailing-newline-empty-blocks": "@general\n@vars\n@all\n@camera\n@options\n@output",
    "grid-types": "@all\n grid = a.b.grid\n@all\n sds = s\n",
    # the fill of per-camera targets / calibration is decided by the LAST @all block alone (upsp_inputs.cpp:76 assigns)
    "all-twice-last-decides": "@all\n grid = g.tri\n targets = all.tgts\n calibration = all.json\n@all\n sds = s.wtd\n"
                              "@camera\n number = 1\n filename = c1.mraw\n",
    "all-twice-last-fills": "@all\n grid = g.tri\n@all\n targets = all.tgts\n@camera\n number = 1\n filename = c1.mraw\n"
                            " calibration = one.json\n@camera\n number = 2\n filename = c2.mraw\n",
}
BROKEN = {
    "bad-registration": "@options\n registration = fancy\n",
    "bad-filter": "@options\n filter = median\n",
    "bad-overlap": "@options\n overlap = all\n",
    "bad-int": "@options\n filter_size = big\n",
    "bad-float": "@options\n oblique_angle = steep\n",
    "bad-run": "@general\n run = x1\n",
    "unresolved-var": "@vars\n a = /x\n@all\n grid = $b/g.tri\n",
    "unresolved-camera-var": "@vars\n a = /x\n@camera\n number = 1\n filename = $zz/v.mraw\n",
}


@pytest.mark.parametrize("name", sorted(DECKS))
def test_deck_reader_equals_reference(probes, tmp_path, name):
    (tmp_path / "d.inp").write_text(DECKS[name])
    mine, ref = both(probes, "deck", tmp_path / "d.inp")
    # accepted with the same fields, or rejected by both (the reference's `$var` substitution passes pos + length as the
    # length of std::string::replace, so a variable in the middle of a value can swallow what follows it: reproduced)
    assert mine[0] == ref[0] and (ref[0] or mine[1] == ref[1]), name
    assert ref[0] == (name in ("vars-twice-and-inside",))


@pytest.mark.parametrize("name", sorted(BROKEN))
def test_deck_reader_rejects_what_the_reference_rejects(probes, tmp_path, name):
    (tmp_path / "d.inp").write_text(BROKEN[name])
    mine, ref = both(probes, "deck", tmp_path / "d.inp")
    assert ref[0] and mine[0], name


def test_deck_check_all_and_writer_equal_reference(probes, tmp_path):
    for n in ("g.tri", "s.wtd", "t.tgts", "c.json", "v.mraw"):
        (tmp_path / n).write_text("x")
    deck = (f"%Version 3.1\n@general\n test = t\n run = 7\n sequence = 2\n tunnel = ames_unitary\n@vars\n d = {tmp_path}\n"
            f"@all\n grid = $d/g.tri\n sds = $d/s.wtd\n targets = $d/t.tgts\n calibration = $d/c.json\n grid_units = in\n"
            f"@camera\n number = 1\n filename = $d/v.mraw\n@options\n filter = gaussian\n filter_size = 3\n@output\n dir = $d\n name = o\n")
    (tmp_path / "ok.inp").write_text(deck)
    mine, ref = both(probes, "deck", tmp_path / "ok.inp", "check")
    assert not ref[0] and mine == ref
    (tmp_path / "bad.inp").write_text(deck.replace("s.wtd", "missing.wtd"))
    mine, ref = both(probes, "deck", tmp_path / "bad.inp", "check")
    assert ref[0] and mine[0]
    # write_file: same text apart from the creation date the reference stamps
    for exe, out in zip(probes, ("mine.inp", "ref.inp")):
        subprocess.run([exe, "deck", str(tmp_path / "ok.inp"), "write", str(tmp_path / out)], check=True, capture_output=True)
    strip = lambda p: [l for l in open(p).read().splitlines() if not l.startswith("%Date_Created")]
    assert strip(tmp_path / "mine.inp") == strip(tmp_path / "ref.inp")


def test_paint_calibration_and_tunnel_conditions_equal_reference(probes, tmp_path):
    (tmp_path / "pc.txt").write_text("a = 1.25\nb=-0.003\n c =  2e-6\nd = 0.75\n e = 0.001\nf = -4.5e-7\ncomment line\ng = 4\na = b = 3\n= 5\n")
    mine, ref = both(probes, "paintcal", tmp_path / "pc.txt", 71.3, 11.82)
    assert not ref[0] and mine == ref
    mine, ref = both(probes, "wtd", os.path.join(GOLDEN, "sample.wtd"))
    assert not ref[0] and mine == ref and len(ref[1]) == 12 and ref[1][-2:] == ["wall_temp 90.2649002", "model_temp 88.125"]
    (tmp_path / "partial.wtd").write_text("RUN 1\n#  MACH Q\tPS   JUNK\n0.5 100.25\t2000  7\n")
    mine, ref = both(probes, "wtd", tmp_path / "partial.wtd")
    assert not ref[0] and mine == ref and "alpha nan" in ref[1]
    path = "/root/reference/test/data/wtd_test.wtd"
    if os.path.exists(path):
        mine, ref = both(probes, "wtd", path)
        assert not ref[0] and mine == ref


def test_model_temperature_equals_reference_code(probes, tmp_path):
    """the model temperature of phase 2 -- recovery-factor wall temperature unless the thermocouple average was read -- from the
    reference's own lines (cpp/exec/psp_process.cpp:2287-2310 with the constants of :1096-1098, compiled from the reference
    tree into ref_probe): subsonic to hypersonic Mach numbers, cold and hot total temperatures, with / without TCAVG, missing
    MACH or TTF (NaN runs through both)."""
    rng = np.random.default_rng(2)
    cases = [(0.0, 70.0, None), (0.84, 97.43, 88.125), (0.84, 97.43, None), (2.5, 150.0, None), (7.0, 900.0, None), (0.3, -40.0, None),
             (1.0, -459.67, None), (None, 97.0, None), (0.6, None, None), (None, None, 75.5)]
    cases += [(float(rng.uniform(0.05, 3.0)), float(rng.uniform(-20, 200)), None if k % 3 else float(rng.uniform(40, 120))) for k in range(50)]
    walls = set()
    for mach, ttf, tc in cases:
        cols = [(n, v) for n, v in (("MACH", mach), ("TTF", ttf), ("TCAVG", tc), ("Q", 650.0), ("PS", 1300.0)) if v is not None]
        (tmp_path / "c.wtd").write_text("RUN 1 1\n#  " + "\t".join(n for n, _ in cols) + "\n" + "\t".join(repr(v) for _, v in cols) + "\n")
        mine, ref = both(probes, "wtd", tmp_path / "c.wtd")
        assert not ref[0] and mine == ref, (mach, ttf, tc)
        d = dict(l.split() for l in ref[1])
        assert d["model_temp"] == (d["tcavg"] if tc is not None else d["wall_temp"])
        walls.add(d["wall_temp"])
    assert len(walls) > 50


def test_plot3d_function_file_and_regression_sample_equal_reference(probes, tmp_path):
    zones = [(4, 5, 1), (3, 4, 1)]
    vals = (np.arange(32, dtype=np.float32) * 0.25 - 3).astype(np.float32)
    for seps in (False, True):                                   # with markers: the reference's one-slot shift, reproduced
        _p3d_function(tmp_path / "f", zones, vals, seps=seps)
        for mode in ((), (1,), (0,)):
            mine, ref = both(probes, "p3dfun", tmp_path / "f", *mode)
            assert mine[0] == ref[0] and (ref[0] or mine[1] == ref[1]), (seps, mode)
    for n, maxels in ((5, 1000), (2500, 1000), (3999, 1000), (12345, 7)):
        (np.arange(n, dtype=np.float32) * np.float32(0.5)).tofile(tmp_path / "v.f32")
        outs = []
        for exe, o in zip(probes, ("m.dat", "r.dat")):
            r = subprocess.run([exe, "vvdump", str(tmp_path / "v.f32"), str(tmp_path / o), str(maxels)], capture_output=True, text=True)
            outs.append((r.stdout, (tmp_path / o).read_bytes()))
        assert outs[0] == outs[1]


def test_plot3d_grid_reader_writer_equal_reference(up, probes, tmp_path):
    base = "/root/reference/cpp/test/inputs/"
    if not os.path.isdir(base):
        pytest.skip("reference fixtures not present on this machine")
    grid_probe = up.build.build_grid_probe()
    for name in sorted(f for f in os.listdir(base) if f.endswith(".x")):
        for prec in ("sp", "dp"):
            # a big-endian file whose precision differs from the grid's: the reference converts the still byte-reversed
            # numbers and swaps afterwards (plot3d.cpp:208-262), which yields garbage; the product swaps first.  Not compared.
            if "bigend" in name and (("_dp" in name) != (prec == "dp")):
                continue
            r1 = subprocess.run([grid_probe, base + name, prec, str(tmp_path / "mine.x")], capture_output=True, text=True)
            r2 = subprocess.run([probes[1], "p3dgrid", base + name, prec, str(tmp_path / "ref.x")], capture_output=True, text=True)
            assert r1.returncode == r2.returncode == 0, (name, r1.stderr, r2.stderr)
            assert r1.stdout == r2.stdout and (tmp_path / "mine.x").read_bytes() == (tmp_path / "ref.x").read_bytes(), (name, prec)


def test_peak_finding_equals_reference(probes, tmp_path):
    """upsp::find_peaks / first_min_threshold (cpp/utils/clustering.ipp:9-101) compiled from the reference tree: histograms
    with plateaus, empty bins (1/0 = inf in the inverse), peaks closer than the separation (the scan's early `break`),
    short and flat inputs."""
    rng = np.random.default_rng(11)
    cases = [np.array(c, np.int32) for c in ([], [3], [1, 2], [1, 5, 2], [5, 5, 5, 5], [0, 0, 0, 0, 0], [1, 3, 3, 3, 1, 0, 0, 4, 4, 2, 9, 1],
                                             [0, 7, 0, 7, 0, 7, 0], [9, 1, 1, 1, 9, 1, 9], [1, 2, 3, 4, 5, 6], [6, 5, 4, 3, 2, 1])]
    for _ in range(40):
        n = int(rng.integers(3, 260))
        base = rng.integers(0, 6, n) * rng.integers(0, 2, n)                 # many zeros and ties
        bumps = (200 * np.exp(-((np.arange(n) - rng.uniform(0, n)) / rng.uniform(2, 20)) ** 2)).astype(int)
        cases.append((base + bumps + (50 * np.exp(-((np.arange(n) - rng.uniform(0, n)) / 6.0) ** 2)).astype(int)).astype(np.int32))
    for k, counts in enumerate(cases):
        counts.tofile(tmp_path / "c.i32")
        for sep in (0, 1, 5, 12):
            mine, ref = both(probes, "peaks", tmp_path / "c.i32", sep)
            assert not ref[0] and mine == ref, (k, sep, counts.tolist())


def test_first_frame_histogram_equals_reference_code(probes, tmp_path):
    """upsp::intensity_histc (cpp/lib/image_processing.ipp:10-50) compiled from the reference tree and called as
    psp_process.cpp:2157 does (bit depth of the camera, 256 bins), followed by first_min_threshold(5): counts, edges, bin size and
    the patcher's threshold equal, for 8- to 16-bit depths (a depth above 16 clamps), pixels at and above the range, bins = -1
    (one bin per value) and other bin counts that divide the range."""
    rng = np.random.default_rng(5)
    imgs = [np.zeros(0, np.uint16), np.zeros(100, np.uint16), np.full(50, 65535, np.uint16), np.arange(65536, dtype=np.uint16),
            np.array([4095, 4096, 4097, 1023, 1024, 255, 256, 0, 15, 16], np.uint16)]
    for seed in range(6):
        lo, hi = rng.uniform(40, 500), rng.uniform(900, 3500)
        imgs.append(np.concatenate([rng.normal(lo, lo / 6, 4000), rng.normal(hi, hi / 8, 50000), rng.normal(30, 8, 300),
                                    rng.integers(0, 65536, 200).astype(float)]).clip(0, 65535).astype(np.uint16))
    n = 0
    for img in imgs:
        img.tofile(tmp_path / "f.u16")
        for depth, bins in [(12, 256), (10, 256), (8, 256), (16, 256), (20, 256), (12, -1), (8, -1), (12, 64), (12, 4096), (10, 1)]:
            mine, ref = both(probes, "hist", tmp_path / "f.u16", depth, bins)
            mine = (mine[0], [l for l in mine[1] if not l.startswith("peaks")])     # the product's probe also lists the maxima
            assert mine == ref and not ref[0], (len(img), depth, bins)
            n += 1
    assert n == 110


def test_area_weighted_node_normals_equal_reference_code(up, probes, tmp_path):
    """upsp::normal / upsp::area of a triangle (cpp/lib/models.ipp:135-184, compiled from the reference tree against the reference's
    own data_structs.h) summed per node as TriModel_::Node::get_normal does (TriModel.ipp:1571-1590): what the camera weights and
    target diameters take as the node normal of an unstructured model == host/grid_readers.hpp, bit for bit, and both == the numpy
    restatement per triangle.  Bumpy spheres, an orphan node, needle / collinear / repeated-corner triangles (Heron's guarded
    term), large offsets, and the reference's own sphere fixtures."""
    from test_grid_readers import numpy_area_weighted_normals, run_probe, write_tri
    rng = np.random.default_rng(3)
    grids = []
    for seed, (nlat, nlon, off) in enumerate([(9, 14, (0.5, -1.0, 2.0)), (6, 9, (1e3, -2e3, 5e2)), (14, 30, (0, 0, 0))]):
        xyz, _, tri = up.synth.make_sphere_mesh(nlat, nlon, 3.0, off, bump=0.15, seed=seed)
        grids.append((np.concatenate([xyz, [[7.0, 7.0, 7.0]]]).astype(np.float32), tri.astype(np.int32)))
    xyz = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1e-7, 0], [5, 5, 5], [5, 5, 5.000001], [1e4, 0, 0], [0, 3, 4], [1, 1, 1]], np.float32)
    tri = np.array([[0, 1, 2], [0, 1, 3], [4, 5, 8], [0, 6, 3], [0, 7, 1], [2, 2, 7], [1, 7, 8], [8, 7, 1], [3, 0, 6]], np.int32)
    grids.append((xyz, tri))
    xyz = (rng.normal(0, 1, (60, 3)) * rng.choice([1e-3, 1.0, 1e3], (60, 1))).astype(np.float32)
    grids.append((xyz, rng.integers(0, 60, (200, 3)).astype(np.int32)))
    base = "/root/reference/cpp/test/inputs/sphere_unf_"
    files = [base + k for k in ("single.tri", "multi.i.tri")] if os.path.exists(base + "single.tri") else []
    for k, (xyz, tri) in enumerate(grids):
        write_tri(tmp_path / f"g{k}.tri", xyz, tri, np.ones(len(tri), np.int32))
        files.append(str(tmp_path / f"g{k}.tri"))
    for k, f in enumerate(files):
        run_probe(up.build.build_grid_probe(), f, tmp_path / "d")
        r = subprocess.run([probes[1], "trigeom", tmp_path / "d.xyz", tmp_path / "d.tri", tmp_path / "ref.f32"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        mine = np.fromfile(tmp_path / "d.nrmw", np.float32).reshape(-1, 3)
        nt = os.path.getsize(tmp_path / "d.tri") // 12
        ref = np.fromfile(tmp_path / "ref.f32", np.float32)
        assert np.array_equal(mine.view(np.uint32), ref[4 * nt:].reshape(-1, 3).view(np.uint32)), f
        xyz = np.fromfile(tmp_path / "d.xyz", np.float32).reshape(-1, 3)
        tri = np.fromfile(tmp_path / "d.tri", np.int32).reshape(-1, 3)
        assert np.array_equal(numpy_area_weighted_normals(xyz, tri).view(np.uint32), mine.view(np.uint32)), f
        assert np.all(np.isfinite(ref)) and np.all(ref[3:4 * nt:4] >= 0)
    assert len(files) >= 5


@pytest.mark.parametrize("n_cams,mode", [(2, "average"), (2, "best"), (3, "best"), (3, "average"), (4, "average")])
def test_camera_weights_equal_reference_code(up, probes, tmp_path, n_cams, mode):
    """the arithmetic of adjust_projection_for_weights from the reference tree: upsp::angle_between (cpp/utils/cv_extras.ipp:67-73)
    and BestView / AverageViews::operator() (cpp/lib/projection.ipp:222-268) applied per multiply-seen node to the cameras in
    ascending order == host/projection_weights.hpp (which visits them in the order the reference's priority queue pops equal rows:
    same container, comparator and push sequence): bit for bit with two cameras and for BestView without exact ties; with >= 3
    averaged cameras the float sum of the angles depends on that order in its last bits (2 ulp allowed, as in
    tests/test_projection_weights.py)."""
    from test_projection_weights import _run, _scene
    xyz, nrm, centers, cams = _scene(up, n_cams, seed=10 + n_cams)
    got, _ = _run(up, tmp_path, xyz, nrm, centers, cams, mode)
    r = subprocess.run([probes[1], "weights", str(tmp_path), str(n_cams), str(len(xyz)), mode], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    changed = 0
    for c in range(n_cams):
        ref = np.fromfile(tmp_path / f"cam{c}.val.ref", np.float32)
        if n_cams == 2 or mode == "best":
            assert np.array_equal(got[c].view(np.uint32), ref.view(np.uint32)), c
        else:
            assert np.all(np.abs(got[c] - ref) <= 2 * np.spacing(np.abs(ref))), c
        changed += int((ref != 1).sum())
    assert changed > 100


@pytest.mark.parametrize("n, F, degree", [(150, 64, 6), (40, 301, 6), (60, 33, 2)])
def test_phase2_node_loop_equals_reference_code(probes, orc, tmp_path, n, F, degree):
    """the per-node loop of the reference's phase 2 -- steady-state pressure -> PaintCalibration::get_gain, Iref / I, detrend,
    delta pressure, delta Cp = p * 144 / qbar, the double rms / avg partial sums, NaN for nodes without coverage -- compiled from
    the reference's own lines (cpp/exec/psp_process.cpp:2460-2498) and linked with its paint calibration and tunnel-condition
    readers; only the least-squares solve inside TransPolyFitter::eval_fit (Eigen) is the oracle's.  orc_phase2, which the GPU
    phase 2 is held against, must give the same bits: delta-Cp histories, rms, avg, gain."""
    rng = np.random.default_rng(n + F)
    t = np.arange(F) / F
    itrans = (900 + 600 * rng.random((n, 1))) * (1 + 0.05 * t[None, :] ** 2 + 0.01 * rng.normal(size=(n, F)))
    itrans = itrans.astype(np.float32)
    itrans[3, 5] = 0.0                                      # a pixel warped in from outside the frame: inf -> NaN history
    avg_final = itrans.astype(np.float64).mean(1).astype(np.float32)
    coverage = (rng.random(n) > 0.1).astype(np.float32) * rng.integers(1, 3, n).astype(np.float32)
    coverage[:4] = [0, 1, 2, 1]
    steady = rng.normal(-0.2, 0.4, n).astype(np.float32)
    temp = rng.uniform(60, 110, n).astype(np.float32)
    cal = np.array([0.62, -1.3e-3, 2.1e-6, 2.4e-4, 3.0e-7, -1.1e-9], np.float32)
    qbar, ps = np.float32(657.9153), np.float32(1332.0421)
    for name, a in (("itrans", itrans), ("avg_final", avg_final), ("coverage", coverage), ("steady", steady), ("model_temp", temp)):
        a.tofile(tmp_path / f"{name}.f32")
    (tmp_path / "paint.cal").write_text("".join("%s = %.9g\n" % (k, v) for k, v in zip("abcdef", cal)))
    (tmp_path / "run.wtd").write_text("RUN 1 1\n#  MACH\tQ\tPS\n0.84\t%.9g\t%.9g\n" % (qbar, ps))
    r = subprocess.run([probes[1], "phase2", str(tmp_path), str(n), str(F), str(degree), orc.build()], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split()[-4:] == ["qbar", "%.9g" % qbar, "ps", "%.9g" % ps]
    pt, rms, avg, gain = orc.phase2(itrans, avg_final, coverage, steady, temp, cal, qbar, ps, degree=degree)
    ref_pt = np.fromfile(tmp_path / "ref_ptrans.f32", np.float32).reshape(n, F)
    with np.errstate(invalid="ignore"):
        ref_rms = np.sqrt(np.fromfile(tmp_path / "ref_rms.f64") / F).astype(np.float32)          # finals, psp_process.cpp:2540-2547
        ref_avg = (np.fromfile(tmp_path / "ref_avg.f64") / F).astype(np.float32)
    ref_gain = np.fromfile(tmp_path / "ref_gain.f64").astype(np.float32)
    same = lambda a, b: np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))
    assert same(pt, ref_pt) and same(rms, ref_rms) and same(avg, ref_avg) and same(gain, ref_gain)
    covered = coverage > 0
    assert np.isnan(gain[~covered]).all() and not ref_pt[~covered].any() and np.isnan(ref_pt[3]).all()
    ok = covered & ~np.isnan(ref_pt).any(1)
    assert ok.sum() > n // 2 and np.all(np.abs(ref_pt[ok]) > 0) and np.all(ref_rms[ok] > 0)


def test_detrend_design_matrix_equals_reference_code(probes, orc, tmp_path):
    """the matrix of the detrend fit, (f / n_frames)^c in the reference's mixed float / double arithmetic: the fill loop of
    TransPolyFitter's constructor (cpp/lib/filtering.ipp:20-24) compiled from the reference tree == orc_transpoly_build, which
    the oracle's phase 2 and the GPU's closed-form detrend are derived from"""
    import ctypes as C
    for F, degree in [(1, 0), (2, 1), (7, 6), (64, 6), (1000, 6), (4097, 6), (20000, 6), (333, 10)]:
        r = subprocess.run([probes[1], "polymat", str(tmp_path / "A.f32"), str(F), str(degree)], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.split() == ["frames", str(F), "coeffs", str(degree + 1)]
        ref = np.fromfile(tmp_path / "A.f32", np.float32)
        A = np.zeros(F * (degree + 1), np.float32)
        orc.lib().orc_transpoly_build(C.c_uint(F), C.c_uint(degree), A.ctypes.data_as(C.c_void_p))
        assert np.array_equal(A.view(np.uint32), ref.view(np.uint32)), (F, degree)
        assert np.all(ref[:F] == 1) and (F < 2 or ref[F + 1] == np.float32(1) / np.float32(F))


def test_finals_equal_reference_code(probes, orc, tmp_path):
    """double partial sums -> the float avg / rms the reference writes, for both phases, and its frame-1 Iref / I - 1 sample
    (cpp/exec/psp_process.cpp:1933-1936, 1947-1949, 2543-2547 compiled from the reference tree) == orc_phase1_finals /
    orc_phase2_finals and the expression tests/test_zz_deck_pipeline.py holds `intensity_ratio_0` to; NaN and zero sums included"""
    import ctypes as C
    rng = np.random.default_rng(9)
    n, F = 500, 3071
    sum_ = rng.uniform(-2e6, 6e6, n)
    sumsq = rng.uniform(0, 4e9, n)
    gain = rng.normal(50, 20, n)
    first = rng.uniform(100, 4000, n).astype(np.float32)
    sum_[:3], sumsq[:3], gain[:3], first[:3] = [np.nan, 0.0, 1e-300], [np.nan, 0.0, 1e-300], [np.nan, 0.0, -0.0], [1000.0, 0.0, 1.0]
    for name, a in (("sum", sum_), ("sumsq", sumsq), ("gain", gain)):
        a.tofile(tmp_path / f"{name}.f64")
    first.tofile(tmp_path / "first.f32")
    r = subprocess.run([probes[1], "finals", str(tmp_path), str(n), str(F)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a1, r1, ratio, a2, r2, g2 = np.fromfile(tmp_path / "ref_finals.f32", np.float32).reshape(6, n)
    avg, rms = orc.phase1_finals(sum_, sumsq, F)
    same = lambda a, b: np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))
    assert same(avg, a1) and same(rms, r1)
    with np.errstate(all="ignore"):
        want = ((avg / first).astype(np.float32).astype(np.float64) - 1.0).astype(np.float32)
    assert same(want, ratio) and np.isnan(ratio[0]) and np.isnan(ratio[1])            # NaN / 1000, 0 / 0
    rf, af, gf = (np.zeros(n, np.float32) for _ in range(3))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    orc.lib().orc_phase2_finals(p(sumsq), p(sum_), p(gain), n, C.c_uint(F), p(rf), p(af), p(gf))
    assert same(af, a2) and same(rf, r2) and same(gf, g2)


def test_seam_remap_equals_reference_code(probes, tmp_path):
    """P3DModel_::adjust_solution (cpp/lib/P3DModel.ipp:146-155, compiled from the reference tree over the reference's
    std::map<node_idx, std::vector<node_idx>>) applied to sol[i] = i == the source-index array of host/p3d_model.hpp, which
    psp_setup_b200 writes as remap.i32 and the GPU applies after every frame (upsp_gpu_set_overlap_remap): the reference's
    3-zone unit-test grid and a sphere of zones with seams, poles (many nodes on one point) and a wrapped zone."""
    from test_p3d_model import reference_fixture, run_probe, write_p3d
    from test_zz_deck_pipeline import sphere_zones
    n_groups = 0
    for k, (zones, tol) in enumerate([(reference_fixture(), 1e-10), (reference_fixture(0.1), 0.100001), (sphere_zones(), 1e-3)]):
        write_p3d(tmp_path / f"g{k}.x", zones)
        info, src, pairs, _, _ = run_probe(probes[0], tmp_path / f"g{k}.x", tol, tmp_path / f"d{k}")
        r = subprocess.run([probes[1], "adjust", str(tmp_path / f"d{k}.pairs"), str(len(src)), str(tmp_path / "sol.f32")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        sol = np.fromfile(tmp_path / "sol.f32", np.float32)
        assert np.array_equal(sol.astype(np.int64), src.astype(np.int64)) and (sol != np.arange(len(src))).sum() > 0, k
        n_groups += int(r.stdout.split()[-1])
    assert n_groups > 20


def test_phase1_accumulation_equals_reference_code(up, probes, orc, tmp_path):
    """the per-frame tail of the reference's phase 1 -- NaN for the nodes no camera sees, then the double partial sums of sol^2
    (float product) and sol -- compiled from its own lines (cpp/exec/psp_process.cpp:1823-1831) and run over the projected
    frames of an oracle case with the NaN marks removed: the marks come back on the same nodes and orc_phase1's sums, which the
    GPU's avg / rms are held against, have the same bits"""
    from chain import Case
    case = Case(up.synth, n_nodes=1500, n_frames=40, height=64, width=96, seed=4)
    inten, s, q = orc.phase1(case.frames, case.csr, first_frame=0, interp=case.interp)
    skipped = np.flatnonzero(np.isnan(inten[0])).astype(np.uint32)
    assert 0 < len(skipped) < case.N // 10 and np.isnan(inten[:, skipped]).all()
    plain = inten.copy()
    plain[:, skipped] = 123.0                       # what project_frame leaves there does not matter: the reference overwrites it
    plain.tofile(tmp_path / "sols.f32")
    skipped.tofile(tmp_path / "skipped.u32")
    r = subprocess.run([probes[1], "accum", str(tmp_path / "sols.f32"), str(case.N), str(tmp_path / "skipped.u32"), str(tmp_path / "o.f64")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split() == ["frames", str(case.F), "nodes", str(case.N), "skipped", str(len(skipped))], r.stderr
    ref_q, ref_s = np.fromfile(tmp_path / "o.f64").reshape(2, case.N)
    keep = ~np.isnan(s)
    assert np.array_equal(np.isnan(ref_s), ~keep) and np.array_equal(np.isnan(ref_q), ~keep) and np.array_equal(np.flatnonzero(~keep), skipped)
    assert np.array_equal(ref_s[keep].view(np.uint64), s[keep].view(np.uint64)) and np.array_equal(ref_q[keep].view(np.uint64), q[keep].view(np.uint64))
    marked = np.fromfile(str(tmp_path / "o.f64") + ".sol", np.float32).reshape(case.F, case.N)
    assert np.array_equal(np.isnan(marked), np.isnan(inten)) and np.array_equal(marked[:, keep], inten[:, keep])


def test_camera_blend_equals_reference_code(up, probes, orc, tmp_path):
    """two weighted cameras on one grid: the reference's own lines that sum the camera solutions of a frame (float, camera
    order; cpp/exec/psp_process.cpp:1813-1819) and mark unseen nodes, fed with each camera's projection alone (the oracle run
    per camera, 0 where that camera sees nothing, as project_frame leaves it) == the oracle's two-camera intensity rows"""
    from chain import Case
    case = Case(up.synth, n_cams=2, n_nodes=1200, n_frames=12, height=64, width=96, seed=7)
    both_, _, _ = orc.phase1(case.frames, case.csr, first_frame=0, interp=case.interp)
    for c in range(2):
        one, _, _ = orc.phase1([case.frames[c]], [case.csr[c]], first_frame=0, interp=case.interp)
        np.nan_to_num(one, nan=0.0).astype(np.float32).tofile(tmp_path / f"cam{c}.sols.f32")
    skipped = np.flatnonzero(np.isnan(both_[0])).astype(np.uint32)
    skipped.tofile(tmp_path / "skipped.u32")
    r = subprocess.run([probes[1], "blend", str(tmp_path), str(case.N), "2", str(tmp_path / "skipped.u32"), str(tmp_path / "o.f32")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = np.fromfile(tmp_path / "o.f32", np.float32).reshape(case.F, case.N)
    seen = ~np.isnan(both_[0])
    assert 0 < len(skipped) < case.N and np.array_equal(np.isnan(ref), np.isnan(both_))
    assert np.array_equal(ref[:, seen].view(np.uint32), both_[:, seen].view(np.uint32))
    a, b = (np.fromfile(tmp_path / f"cam{c}.sols.f32", np.float32).reshape(case.F, case.N) for c in range(2))
    assert ((a != 0) & (b != 0)).any(0).sum() > 50                       # nodes both cameras contribute to


def test_unpack_restatement_equals_reference_code(probes, orc, tmp_path):
    """a1: the reference's own upsp::unpack_12bit / unpack_10bit (cpp/lib/PSPVideo.cpp:111-150, compiled from the reference
    tree) against the CPU restatement the GPU decoder is held to, on random packed bytes (every bit pattern class)."""
    rng = np.random.default_rng(21)
    for bits, group in ((12, 3), (10, 5)):
        npix = 4096 * 3
        packed = rng.integers(0, 256, npix * bits // 8, dtype=np.uint8)
        packed[: group * 4] = [0xFF] * (group * 2) + [0x00] * (group * 2)
        packed.tofile(tmp_path / "p.bin")
        r = subprocess.run([probes[1], "unpack", str(tmp_path / "p.bin"), str(bits), str(npix), str(tmp_path / "o.u16")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        ref = np.fromfile(tmp_path / "o.u16", np.uint16)
        mine = orc.unpack_12bit(packed) if bits == 12 else orc.unpack_10bit(packed)
        assert np.array_equal(mine.ravel(), ref)


def test_mraw_reader_equals_reference_code(up, probes, orc, tmp_path):
    """the reference's own MrawReader + PSPVideo (cpp/lib/MrawReader.cpp:62-146) on the committed .mraw/.cih pair, on a
    generated one and on the reference's 12bitMRAW fixture: same properties as host/video_readers.hpp reports, and its decoded
    frames == unpack(the stored bytes the product hands to the GPU)."""
    video_probe = up.build.build_probe()
    files = [os.path.join(GOLDEN, "tiny12.mraw")]
    rng = np.random.default_rng(4)
    W, H, F = 48, 20, 5
    fr = rng.integers(0, 4096, (F, H * W)).astype(np.uint16)
    up.synth.pack_12bit(fr).tofile(tmp_path / "gen.mraw")
    (tmp_path / "gen.cih").write_text("#Camera Information Header\r\nRecord Rate(fps) : 2500\r\nTotal Frame : %d\r\n"
                                      "Image Width : %d\r\nImage Height : %d\r\nColor Bit : 12\r\n" % (F, W, H))
    files.append(str(tmp_path / "gen.mraw"))
    if os.path.exists("/root/reference/cpp/test/mraw/12bitMRAW.mraw"):
        files.append("/root/reference/cpp/test/mraw/12bitMRAW.mraw")
    for path in files:
        r = subprocess.run([probes[1], "mraw", path, "1", "2", str(tmp_path / "ref.u16")], capture_output=True, text=True)
        m = subprocess.run([video_probe, path, "1", "2", str(tmp_path / "mine.bin")], capture_output=True, text=True)
        assert r.returncode == 0 and m.returncode == 0, (r.stderr, m.stderr)
        rk = dict(l.split() for l in r.stdout.splitlines())
        mk = dict(l.split()[:2] for l in m.stdout.splitlines() if not l.startswith("crc"))
        for k in ("width", "height", "bit_depth", "num_frames"):
            assert rk[k] == mk[k], (path, k)
        assert float(rk["frame_rate"]) == float(mk["frame_rate"])
        decoded = np.fromfile(tmp_path / "ref.u16", np.uint16)
        stored = np.fromfile(tmp_path / "mine.bin", np.uint8)
        assert np.array_equal(orc.unpack_12bit(stored).ravel(), decoded), path
    assert np.array_equal(np.fromfile(tmp_path / "ref.u16", np.uint16).size, 2 * int(rk["width"]) * int(rk["height"]))


def test_hot_pixel_restatement_equals_reference_code(probes, orc, tmp_path):
    """a2: the reference's own upsp::fix_hot_pixels (cpp/utils/cv_extras.cpp:230-272, compiled from the reference tree)
    against the CPU restatement the GPU path is held to: 0 / 1 / 5 / 6 hot pixels, neighbours that are hot themselves,
    corners and borders, the >= 4064 and > 512 boundaries, fixes that feed later fixes (raster order), random frames."""
    rng = np.random.default_rng(31)
    H, W = 24, 40
    frames = []

    def frame(base=1500):
        return rng.integers(base - 200, base + 200, (H, W)).astype(np.uint16)
    frames.append(frame())                                            # no hot pixel
    f = frame(); f[5, 7] = 4095; frames.append(f)                     # one
    f = frame(); f[[0, 0, H - 1, H - 1, 9], [0, W - 1, 0, W - 1, 0]] = 4095; frames.append(f)    # five: corners + border
    f = frame(); f[3, 3:9] = 4095; frames.append(f)                   # six in a row: too many, untouched
    f = frame(); f[8, 8] = 4095; f[8, 9] = 4080; f[9, 8] = 4064; frames.append(f)       # hot neighbours, raster order
    f = frame(); f[4, 4] = 4063; f[6, 6] = 4064; frames.append(f)     # threshold boundary
    f = np.full((H, W), 3552, np.uint16); f[10, 10] = 4064; frames.append(f)            # old - new == 512: kept
    f = np.full((H, W), 3551, np.uint16); f[10, 10] = 4064; frames.append(f)            # old - new == 513: replaced
    f = frame(3900); f[2, 2] = 4070; f[2, 3] = 4069; frames.append(f)                   # change too small
    f = frame(); f[0, 5] = 4095; f[1, 5] = 4095; f[H - 1, 20] = 4095; f[12, W - 1] = 4095; frames.append(f)   # 2/3-neighbour medians
    for _ in range(30):
        f = frame(int(rng.integers(500, 3800)))
        n = int(rng.integers(0, 8))
        f[rng.integers(0, H, n), rng.integers(0, W, n)] = rng.integers(4000, 4096, n)
        frames.append(f)
    stack = np.stack(frames)
    stack.tofile(tmp_path / "in.u16")
    r = subprocess.run([probes[1], "hotpix", str(tmp_path / "in.u16"), str(H), str(W), str(tmp_path / "out.u16"), str(len(frames))],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = np.fromfile(tmp_path / "out.u16", np.uint16).reshape(stack.shape)
    changed = 0
    for k, f in enumerate(frames):
        mine, _ = orc.fix_hot_pixels(f)
        assert np.array_equal(mine, ref[k]), k
        changed += int((ref[k] != f).sum())
    assert changed >= 12 and np.array_equal(ref[3], frames[3]) and ref[6][10, 10] == 4064 and ref[7][10, 10] == 3551


def test_ray_cast_restatement_equals_reference_code(up, probes, orc, tmp_path):
    """Visibility of create_projection_mat: the reference's own ray caster (cpp/raycast/pspRT.cpp: SAH BVH build, its traversal
    and the watertight rt::Triangle::intersect, compiled from the reference tree; createBVH's call sequence psp_process.cpp:45-53)
    against the restatement the GPU operator is held to (nearest hit over all triangles).  Rays as psp_process shoots them:
    camera centre -> every node (normalised, :257-259), the six +-1e-4 jittered retries (un-normalised, :273-275), plus rays
    that miss, rays along axes and rays through shared edges / vertices (ties)."""
    sc = up.synth.make_projection_scene(n_lat=20, n_lon=40, seed=5)
    xyz, tri = sc["xyz"].astype(np.float32), sc["tri"].astype(np.int32)
    ocam = orc.make_camera(sc["rvec"], sc["tvec"], sc["K"], sc["dist"], sc["width"], sc["height"])
    orig = orc.cam_center(ocam).astype(np.float32)
    rng = np.random.default_rng(8)
    d = (xyz - orig).astype(np.float32)
    ln = np.sqrt((d * d).sum(1, dtype=np.float32)).astype(np.float32)
    rays = [np.concatenate([np.tile(orig, (len(xyz), 1)), (d / ln[:, None]).astype(np.float32)], 1)]
    for axis in range(3):
        for s in (-1e-4, 1e-4):
            p = xyz.copy()
            p[:, axis] += np.float32(s)
            rays.append(np.concatenate([np.tile(orig, (len(xyz), 1)), (p - orig).astype(np.float32)], 1)[::7])
    mid = ((xyz[tri[:, 0]] + xyz[tri[:, 1]]) * np.float32(0.5)).astype(np.float32)          # through shared edges
    rays.append(np.concatenate([np.tile(orig, (len(mid), 1)), (mid - orig).astype(np.float32)], 1)[::5])
    other = np.float32([40.0, 3.0, 1.0])                                                     # a second viewpoint, random directions
    rd = rng.normal(0, 1, (400, 3)).astype(np.float32)
    rays.append(np.concatenate([np.tile(other, (400, 1)), rd], 1))
    rays.append(np.float32([[0, 0, -30, 0, 0, 1], [0, 0, -30, 0, 0, -1], [0, 0, -30, 1, 0, 0], [0, -30, 0, 0, 1, 0], [30, 0, 0, -1, 0, 0]]))
    rays = np.ascontiguousarray(np.concatenate(rays), np.float32)
    xyz[tri].reshape(-1, 9).astype(np.float32).tofile(tmp_path / "tris.f32")                 # extract_tris layout
    rays.tofile(tmp_path / "rays.f32")
    r = subprocess.run([probes[1], "raycast", str(tmp_path / "tris.f32"), str(tmp_path / "rays.f32"), str(tmp_path / "out.bin")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(tmp_path / "out.bin", np.uint8).reshape(-1, 12)
    ref_hit, ref_t, ref_prim = raw[:, 0:4].copy().view(np.int32)[:, 0], raw[:, 4:8].copy().view(np.float32)[:, 0], raw[:, 8:12].copy().view(np.int32)[:, 0]
    hit, t, prim = orc.cast_rays(xyz, tri, rays)
    assert np.array_equal(hit, ref_hit) and 0.3 < hit.mean() < 0.999
    h = hit == 1
    assert np.array_equal(t[h].view(np.uint32), ref_t[h].view(np.uint32))                    # nearest distance, bit for bit
    same = prim[h] == ref_prim[h]
    # a different triangle may only be reported where several triangles are hit at exactly that distance (shared edge /
    # vertex): the reference keeps the first one its traversal meets, the restatement the lowest index
    if not same.all():
        idx = np.flatnonzero(h)[~same]
        for i in idx[:50]:
            _, t_other, _ = orc.cast_rays(xyz, tri[[ref_prim[i]]], rays[[i]])
            assert t_other[0] == t[i]
    assert same.mean() > 0.9
    print("rays %d, hits %d, same triangle %d, tie-different triangle %d" % (len(rays), h.sum(), same.sum(), (~same).sum()))


def _kdtree(probes, tmp_path, pts, tol, queries):
    pts.astype(np.float32).tofile(tmp_path / "pts.f32")
    queries.astype(np.float32).tofile(tmp_path / "q.f32")
    r = subprocess.run([probes[1], "kdtree", str(tmp_path / "pts.f32"), repr(float(tol)), str(tmp_path / "q.f32"), str(tmp_path / "kd.txt")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    ranges, nearest = [], []
    for line in open(tmp_path / "kd.txt"):
        head, _, rest = line.partition(":")
        (ranges if head.startswith("range") else nearest).append([int(v) for v in rest.split()])
    return ranges, [n[0] for n in nearest]


def test_kdtree_queries_equal_reference_code(probes, tmp_path):
    """The reference's kd-tree (cpp/raycast/pspKdtree.c, compiled as C from the reference tree) as the models use it:
    kd_nearest_range3 around every zone-edge node (P3DModel_::identifyOverlap, P3DModel.ipp:954-1003) returns exactly the nodes
    whose squared distance in double is <= tol^2 -- the predicate host/p3d_model.hpp sweeps with -- including the reference's
    own tolerance cases (offset 0.1 against tol 0.09 / 0.100001, test_p3dmodel.cpp:205-236); kd_nearest (getTargets,
    get_target_diameters) returns the closest node, one of the closest on exact ties."""
    from test_p3d_model import reference_fixture, seam_grid
    rng = np.random.default_rng(3)
    cases = [(np.concatenate([z[2] for z in reference_fixture(0.1)]), 0.09), (np.concatenate([z[2] for z in reference_fixture(0.1)]), 0.100001),
             (np.concatenate([z[2] for z in reference_fixture(0.0)]), 1e-10), (np.concatenate([z[2] for z in seam_grid(1)]), 1e-3)]
    path = "/root/reference/test/data/fml_tc3_volume.grid"
    if os.path.exists(path):
        raw = open(path, "rb").read()
        nz = struct.unpack_from("<i", raw, 4)[0]
        dims = np.frombuffer(raw, "<i4", 3 * nz, 16).reshape(nz, 3)
        off, edge = 16 + 12 * nz + 4, []
        for j, k, l in dims:
            n = int(j * k * l)
            p = np.frombuffer(raw, "<f4", 3 * n, off + 4).reshape(3, int(k), int(j)).transpose(1, 2, 0)
            edge += [p[0], p[-1], p[:, 0], p[:, -1]]
            off += 12 * n + 8
        cases.append((np.concatenate(edge)[::3], 1e-3))                  # a third of the zone-edge nodes of the 14-zone grid
    for pts, tol in cases:
        pts = np.ascontiguousarray(pts, np.float32)
        q = np.concatenate([pts[rng.choice(len(pts), 40)] + rng.normal(0, 0.3, (40, 3)), pts[:5]]).astype(np.float32)
        ranges, nearest = _kdtree(probes, tmp_path, pts, tol, q)
        t = float(np.float32(tol))
        P = pts.astype(np.float64)
        for i in range(0, len(pts), max(1, len(pts) // 1500)):           # every node for the small cases, a sample otherwise
            want = np.flatnonzero(((P - P[i]) ** 2).sum(1) <= t * t)
            assert ranges[i] == want.tolist(), (tol, i)
        for k, qq in enumerate(q.astype(np.float64)):
            d2 = ((P - qq) ** 2).sum(1)
            assert d2[nearest[k]] == d2.min()


def test_patch_geometry_equals_reference_code(up, probes, tmp_path):
    """Which pixels get patched: the reference's own cluster_points / get_target_boundary / get_cluster_boundary /
    PatchClusters constructor / threshold_bounds (cpp/lib/patches.ipp:15-94, 240-487, header templates compiled from the
    reference tree) against host/patch_geometry.hpp, through the two probes' identical text output: single targets, touching
    and chained targets (multi-target clusters), targets at and beyond the frame border, several thickness settings, with and
    without the first-frame threshold."""
    mine = up.build.build_patch_probe()
    rng = np.random.default_rng(17)
    W, H = 160, 120
    ref = rng.integers(800, 3000, (H, W)).astype(np.uint16)
    for _ in range(60):                                                   # dark blobs: boundary pixels near them are dropped
        y, x = int(rng.integers(0, H)), int(rng.integers(0, W))
        ref[max(0, y - 1):y + 2, max(0, x - 1):x + 2] = rng.integers(0, 300)
    ref.tofile(tmp_path / "ref.u16")
    sets = []
    for k in range(12):
        n = int(rng.integers(1, 14))
        t = np.stack([rng.uniform(-3, W + 3, n), rng.uniform(-3, H + 3, n), rng.uniform(2.0, 9.0, n)], 1)
        if k % 3 == 0:                                                    # a chain of touching targets and a tight pair
            t = np.concatenate([t, [[40 + 7 * i, 60 + 2 * i, 6.0] for i in range(5)], [[100.5, 20.25, 4.0], [103.0, 22.5, 5.5]]])
        if k % 4 == 1:
            t = np.concatenate([t, [[0.0, 0.0, 5.0], [W - 1.0, H - 1.0, 7.0], [W / 2, 0.4, 3.0], [-6.0, 50.0, 4.0]]])
        sets.append(t.astype(np.float32))
    n_multi = 0
    for k, t in enumerate(sets):
        (tmp_path / "t.txt").write_text("".join("%.9g %.9g %.9g\n" % tuple(float(v) for v in row) for row in t))
        for bt, bf in ((2, 1), (2, 0), (1, 2), (3, 1)):
            for thr in ((), (str(tmp_path / "ref.u16"), "400", "2"), (str(tmp_path / "ref.u16"), "1200", "1")):
                args = [str(tmp_path / "t.txt"), str(W), str(H), str(bt), str(bf)] + list(thr)
                a = subprocess.run([mine] + args, capture_output=True, text=True)
                b = subprocess.run([probes[1], "patches"] + args, capture_output=True, text=True)
                assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
                assert a.stdout == b.stdout, (k, bt, bf, thr)
        n_multi += sum(1 for l in b.stdout.splitlines() if l.startswith("cluster") and int(l.split()[2]) > 1)
    assert n_multi >= 4


def test_apportion_equals_reference_code(probes, orc, up):
    """the work split of frames and nodes over ranks (apportion, psp_process.cpp:611-624; the identical function of the
    stand-alone transpose tool is what is linked here): the restatement and the CUDA library's slices follow it."""
    cases = [(20000, 1), (20000, 8), (50000, 8), (1000000, 8), (500000, 3), (7, 8), (8, 8), (9, 8), (0, 4), (1, 1), (12345, 7), (100, 16)]
    for value, nbins in cases:
        r = subprocess.run([probes[1], "apportion", str(value), str(nbins)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        ref = {l.split()[0]: [int(v) for v in l.split()[1:]] for l in r.stdout.splitlines()}
        start, extent = orc.apportion(value, nbins)
        assert list(map(int, start)) == ref["start"] and list(map(int, extent)) == ref["extent"], (value, nbins)
