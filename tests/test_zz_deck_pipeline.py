"""The reference's own inputs end to end on the GPU (file name sorts last on purpose: every other parity test runs
first): input deck + .tri grid + calibration JSON + Photron .mraw + paint calibration + .wtd
  -> psp_setup_b200 -input_file (projection matrix on the GPU, job directory)
  -> psp_process_b200 -job_dir  (frame chain, the reference's flat files)
held against the oracle: create_projection_mat restated on the same grid / calibration, then the frame chain on the
decoded frames with the deck's run constants."""
import os
import subprocess

import numpy as np
import pytest

from chain import Case, run_oracle, same_bits
from conftest import GOLDEN
from test_grid_readers import write_tri
from test_setup_tool import _cal_json


@pytest.mark.gpu
def test_deck_to_flat_files(up, orc, gpu, tmp_path):
    import cv2
    synth = up.synth
    sc = synth.make_projection_scene(n_lat=24, n_lon=48, seed=11)
    W, H, F = sc["width"], sc["height"], 12
    d = tmp_path
    write_tri(d / "model.tri", sc["xyz"], sc["tri"], np.ones(len(sc["tri"]), np.int32))
    _cal_json(d / "cam01.json", cv2.Rodrigues(np.asarray(sc["rvec"], float))[0], sc["tvec"], sc["K"], sc["dist"], (W, H))
    frames = synth.make_frames(F, H, W, seed=5)[0]
    synth.pack_12bit(frames.reshape(F, -1)).tofile(d / "v1.mraw")
    (d / "v1.cih").write_text("#Camera Information Header\r\nRecord Rate(fps) : 1000\r\nTotal Frame : %d\r\n"
                              "Image Width : %d\r\nImage Height : %d\r\nColor Bit : 12\r\n" % (F, W, H))
    (d / "run.wtd").write_text(open(os.path.join(GOLDEN, "sample.wtd")).read())
    (d / "model.tgts").write_text(open(os.path.join(GOLDEN, "sample.tgts")).read())
    cal = np.array([0.62, -1.3e-3, 2.1e-6, 2.4e-4, 3.0e-7, -1.1e-9], np.float32)
    (d / "paint.cal").write_text("".join("%s = %.9g\n" % (k, v) for k, v in zip("abcdef", cal)))
    (d / "out").mkdir()
    (d / "job").mkdir()
    (d / "deck.inp").write_text(
        f"%Version 0.0\n@general\n\ttest = t-b200\n\trun = 12\n\tsequence = 3\n\ttunnel = ames_unitary\n@vars\n\td = {d}\n"
        f"@all\n\tgrid = $d/model.tri\n\tsds = $d/run.wtd\n\ttargets = $d/model.tgts\n"
        f"@camera\n\tnumber = 1\n\tfilename = $d/v1.mraw\n\tcalibration = $d/cam01.json\n"
        f"@options\n\ttarget_patcher = none\n\tregistration = none\n\tfilter = none\n\tfilter_size = 1\n\toblique_angle = 70\n"
        f"\tnumber_frames = -1\n@output\n\tdir = $d/out\n\tname = run12\n")
    # one step, the reference's own command line (launcher style -key=value, python/upsp/processing/tree.py:455-465)
    r = subprocess.run([up.build.build_host(), f"-input_file={d / 'deck.inp'}", f"-h5_out={d / 'out' / 'run12.h5'}",
                        f"-paint_cal={d / 'paint.cal'}", f"-add_out_dir={d / 'out'}", "-frames=12", "-chunk", "8"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    jobd = d / "out" / "job_b200"

    # the projection matrix the oracle builds from the same grid and calibration
    subprocess.run([up.build.build_grid_probe(), str(d / "model.tri"), str(d / "g")], check=True, capture_output=True)
    nrm = np.fromfile(d / "g.nrm", np.float32).reshape(-1, 3)
    pc = subprocess.run([up.build.build_setup_tool(), "-cal", str(d / "cam01.json"), "-print_cal"], capture_output=True, text=True)
    parsed = {l.split()[0]: np.array(l.split()[1:], float) for l in pc.stdout.splitlines()}
    ocam = orc.make_camera(parsed["rvec"], parsed["tvec"], sc["K"], sc["dist"], W, H)
    thresh = np.float32((180.0 - 70.0) * np.pi / 180.0)
    N = len(sc["xyz"])
    code, uv = orc.create_projection(ocam, sc["xyz"], nrm, np.ones(N, np.uint8), sc["tri"], float(thresh))
    rowptr, col, val = orc.projection_csr(code)
    assert (code >= 0).sum() > 100
    assert np.array_equal(np.fromfile(jobd / "cam0.rowptr", np.int32), rowptr)
    assert np.array_equal(np.fromfile(jobd / "cam0.col", np.int32), col)
    assert np.array_equal(np.fromfile(jobd / "cam0.val", np.float32), val)          # one camera: weights stay 1
    assert np.array_equal(np.fromfile(d / "out" / "cam01-uv", np.float32).view(np.uint32), uv.ravel().view(np.uint32))

    case = Case.__new__(Case)
    case.C, case.N, case.F, case.H, case.W = 1, N, F, H, W
    case.interp, case.degree, case.fmt, case.filter_kind, case.filter_size = 1, 6, "p12", 0, 0
    case.frames, case.csr, case.warp, case.patch_lists, case.overlap, case.synth = [frames], [(rowptr, col, val)], None, None, None, synth
    case.cal, case.qbar, case.ps = cal, np.float32(657.9153), np.float32(1332.0421)
    case.steady, case.temp = np.zeros(N, np.float32), np.full(N, 88.125, np.float32)
    ref = run_oracle(orc, case)
    rd = lambda n, shape=None: (np.fromfile(d / "out" / n, np.float32).reshape(shape) if shape else np.fromfile(d / "out" / n, np.float32))
    it = rd("intensity_transpose", (N, F))
    assert same_bits(it, ref["itrans"]) and same_bits(rd("intensity_avg"), ref["avg"]) and same_bits(rd("intensity_rms"), ref["rms"])
    assert same_bits(rd("coverage"), ref["coverage"]) and same_bits(rd("gain"), ref["gain"])
    xyz = sc["xyz"].astype(np.float32)
    assert np.array_equal(rd("X"), xyz[:, 0]) and np.array_equal(rd("Y"), xyz[:, 1]) and np.array_equal(rd("Z"), xyz[:, 2])
    with np.errstate(all="ignore"):
        ratio0 = ((ref["avg"] / ref["itrans"][:, 0]).astype(np.float32).astype(np.float64) - 1.0).astype(np.float32)
    assert same_bits(rd("intensity_ratio_0"), ratio0)
    assert same_bits(rd("model_temp"), case.temp) and np.array_equal(rd("steady_state"), case.steady)
    # regression samples (psp_process.cpp:2006-2015): N = 1284 < 2000 -> stride 1, the first 1000 values
    assert same_bits(rd("vv-int-avg.dat"), ref["avg"][:1000]) and same_bits(rd("vv-int-sample1.dat"), ratio0[:1000])
    assert same_bits(rd("vv-cp-rms.dat"), rd("rms")[:1000])
    # the two HDF5 files (psp_process.cpp:2400-2420, 2535-2604), written by host/psp_hdf5.hpp and read by tests/h5min.py
    import h5min
    h5, ex = h5min.File(str(d / "out" / "run12.h5")), h5min.File(str(d / "out" / "extras.h5"))
    for f, transposed in ((h5, 1), (ex, 0)):
        assert f.root.attrs["psph5_version"][0] == 1 and f.root.attrs["nodal"][0] == 1 and f.root.attrs["structured"][0] == 0
        assert f.root.attrs["transpose"][0] == transposed
        assert np.array_equal(f.root["Grid/x"].data, xyz[:, 0]) and np.array_equal(f.root["Grid/z"].data, xyz[:, 2])
        assert np.array_equal(f.root["Grid/triangles"].data, np.asarray(sc["tri"], np.uint32).reshape(-1, 3))
        assert same_bits(f.root["rms"].data, rd("rms")) and same_bits(f.root["coverage"].data, ref["coverage"])
        assert same_bits(f.root["model_temp"].data, case.temp) and f.root["rms"].attrs["units"] == ["delta Cp"]
        assert f.root["Condition/dynamic_pressure"].data[0] == case.qbar and f.root["Condition/static_pressure"].data[0] == case.ps
        assert f.root["Condition/frame_rate"].attrs["units"] == ["Hz"] and f.root["Condition/focal_length"].data.shape == (1,)
    assert same_bits(ex.root["average"].data, rd("avg")) and "average" not in h5.root.children
    # the documented last step of a run (docs/sphinx/quick-start.rst:125-160): add_field <h5> frames <pressure_transpose> <frames>
    r = subprocess.run([up.build.build_add_field_tool(), str(d / "out" / "run12.h5"), "frames", str(d / "out" / "pressure_transpose"),
                        str(F)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    h5b = h5min.File(str(d / "out" / "run12.h5"))
    assert h5b.root["frames"].shape == (N, F) and same_bits(h5b.root["frames"].data, rd("pressure_transpose", (N, F)))
    assert same_bits(h5b.root["rms"].data, rd("rms")) and np.array_equal(h5b.root["Grid/x"].data, xyz[:, 0])


@pytest.mark.gpu
def test_deck_with_polynomial_patcher(up, orc, gpu, tmp_path):
    """deck target_patcher = polynomial: .tgts -> visible / projected / sized targets (host/targets.hpp, setup tool) ->
    clusters, histogram threshold on the first frame decoded from the video, pixel lists (driver) -> patched chain."""
    import cv2
    from oracle import setup_patches as sp
    synth = up.synth
    sc = synth.make_projection_scene(n_lat=24, n_lon=48, seed=11)
    W, H, F = sc["width"], sc["height"], 10
    d = tmp_path
    write_tri(d / "model.tri", sc["xyz"], sc["tri"], np.ones(len(sc["tri"]), np.int32))
    _cal_json(d / "cam01.json", cv2.Rodrigues(np.asarray(sc["rvec"], float))[0], sc["tvec"], sc["K"], sc["dist"], (W, H))
    frames = synth.make_frames(F, H, W, seed=6)[0]
    synth.pack_12bit(frames.reshape(F, -1)).tofile(d / "v1.mraw")
    (d / "v1.cih").write_text("#Camera Information Header\r\nRecord Rate(fps) : 1000\r\nTotal Frame : %d\r\n"
                              "Image Width : %d\r\nImage Height : %d\r\nColor Bit : 12\r\n" % (F, W, H))
    (d / "run.wtd").write_text(open(os.path.join(GOLDEN, "sample.wtd")).read())
    # fiducials painted on visible grid nodes (so that nodes do read patched pixels) + one on the far side
    subprocess.run([up.build.build_grid_probe(), str(d / "model.tri"), str(d / "g")], check=True, capture_output=True)
    nrm = np.fromfile(d / "g.nrm", np.float32).reshape(-1, 3)
    ocam = orc.make_camera(sc["rvec"], sc["tvec"], sc["K"], sc["dist"], W, H)
    code, _ = orc.create_projection(ocam, sc["xyz"], nrm, np.ones(len(sc["xyz"]), np.uint8), sc["tri"], float(np.float32(110 * np.pi / 180)))
    vis = np.nonzero(code >= 0)[0]
    inner = [n for n in vis if 40 < code[n] % W < W - 40 and 40 < code[n] // W < H - 40]
    picks = [inner[i] for i in np.linspace(0, len(inner) - 1, 6).astype(int)]
    pts = [sc["xyz"][n] for n in picks] + [np.array([0.0, 0.0, 5.0])]
    rows = []
    for i, p in enumerate(pts):
        rows.append("%4d %10.4f %10.4f %10.4f 0.0 0.0 1.0 %6.3f 1 2 3 st%02d\n" % (i + 1, p[0], p[1], p[2], 0.05 + 0.004 * i, i + 1))
    (d / "model.tgts").write_text("*Targets\n" + "".join(rows[:5]) + "*Fiducials\n" + "".join(rows[5:]))
    cal = np.array([0.62, -1.3e-3, 2.1e-6, 2.4e-4, 3.0e-7, -1.1e-9], np.float32)
    (d / "paint.cal").write_text("".join("%s = %.9g\n" % (k, v) for k, v in zip("abcdef", cal)))
    (d / "out").mkdir()
    (d / "job").mkdir()
    (d / "deck.inp").write_text(
        f"@general\n\ttest = t\n\trun = 1\n\tsequence = 1\n\ttunnel = ames_unitary\n@all\n\tgrid = {d}/model.tri\n\tsds = {d}/run.wtd\n"
        f"\ttargets = {d}/model.tgts\n@camera\n\tnumber = 1\n\tfilename = {d}/v1.mraw\n\tcalibration = {d}/cam01.json\n"
        f"@options\n\ttarget_patcher = polynomial\n\tregistration = none\n\tfilter = none\n\tfilter_size = 1\n\toblique_angle = 70\n"
        f"\tnumber_frames = {F}\n@output\n\tdir = {d}/out\n\tname = r\n")
    r = subprocess.run([up.build.build_setup_tool(), "-input_file", str(d / "deck.inp"), "-paint_cal", str(d / "paint.cal"),
                        "-job_dir", str(d / "job")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    targs = [tuple(np.float32(x) for x in row) for row in np.loadtxt(d / "job" / "cam0.targets", dtype=np.float32).reshape(-1, 3)]
    assert len(targs) == 6                                       # the far-side target is hidden
    r = subprocess.run([up.build.build_host(), "-job_dir", str(d / "job"), "-out_dir", str(d / "out"), "-chunk", "8"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert f"Sorted {len(targs)} targets into" in r.stdout

    first = frames[0]              # cams[c]->get_frame(1) as InitializeImagePatches takes it: no hot-pixel fix (psp_process.cpp:2091)
    first.astype("<u2").tofile(d / "first.u16")
    hp = subprocess.run([up.build.build_inputs_probe(), "hist", str(d / "first.u16"), "12"], capture_output=True, text=True)
    thresh = int([l for l in hp.stdout.splitlines() if l.startswith("threshold")][0].split()[1])
    bt, bf = 2, 1
    lists = sp.patch_clusters(sp.cluster_points(targs, bt + bf), W, H, bt, bf, first, thresh, 2)
    xy = lambda pts: (np.array([p[0] for p in pts], np.uint32), np.array([p[1] for p in pts], np.uint32))
    N = len(sc["xyz"])
    case = Case.__new__(Case)
    case.C, case.N, case.F, case.H, case.W = 1, N, F, H, W
    case.interp, case.degree, case.fmt, case.filter_kind, case.filter_size = 1, 6, "p12", 0, 0
    csr = tuple(np.fromfile(d / "job" / ("cam0." + k), t) for k, t in (("rowptr", np.int32), ("col", np.int32), ("val", np.float32)))
    case.frames, case.csr, case.warp, case.overlap, case.synth = [frames], [csr], None, None, synth
    case.patch_lists = [([xy(b) for b, _ in lists], [xy(i) for _, i in lists])]
    case.cal, case.qbar, case.ps = cal, np.float32(657.9153), np.float32(1332.0421)
    case.steady, case.temp = np.zeros(N, np.float32), np.full(N, 88.125, np.float32)
    ref = run_oracle(orc, case)
    got = np.fromfile(d / "out" / "intensity_transpose", np.float32).reshape(N, F)
    assert same_bits(got, ref["itrans"])
    assert same_bits(np.fromfile(d / "out" / "intensity_avg", np.float32), ref["avg"])
    plain = Case.__new__(Case)
    plain.__dict__.update(case.__dict__)
    plain.patch_lists = None
    assert not same_bits(got, run_oracle(orc, plain)["itrans"])   # the patches did touch pixels that nodes read


@pytest.mark.gpu
def test_deck_two_cameras_average_view(up, orc, gpu, tmp_path):
    """two cameras on one model: per-camera projection matrices, adjust_projection_for_weights with the camera centres
    (psp_process.cpp:1623-1641; AverageViews), the chain sums the weighted camera solutions (:1814-1820)."""
    import cv2
    from test_projection_weights import _angles
    synth = up.synth
    sc = synth.make_projection_scene(n_lat=24, n_lon=48, seed=11)
    W, H, F = sc["width"], sc["height"], 8
    d = tmp_path
    write_tri(d / "model.tri", sc["xyz"], sc["tri"], np.ones(len(sc["tri"]), np.int32))
    rvecs = [np.asarray(sc["rvec"], float), np.array([0.05, 0.33, 0.3])]
    frames = []
    cams_txt = ""
    for c, rv in enumerate(rvecs):
        _cal_json(d / f"cam{c + 1:02d}.json", cv2.Rodrigues(rv)[0], sc["tvec"], sc["K"], sc["dist"], (W, H))
        frames.append(synth.make_frames(F, H, W, seed=7 + c)[0])
        synth.pack_12bit(frames[c].reshape(F, -1)).tofile(d / f"v{c + 1}.mraw")
        (d / f"v{c + 1}.cih").write_text("#Camera Information Header\r\nRecord Rate(fps) : 1000\r\nTotal Frame : %d\r\n"
                                         "Image Width : %d\r\nImage Height : %d\r\nColor Bit : 12\r\n" % (F, W, H))
        cams_txt += f"@camera\n\tnumber = {c + 1}\n\tfilename = {d}/v{c + 1}.mraw\n\tcalibration = {d}/cam{c + 1:02d}.json\n"
    (d / "run.wtd").write_text(open(os.path.join(GOLDEN, "sample.wtd")).read())
    (d / "model.tgts").write_text(open(os.path.join(GOLDEN, "sample.tgts")).read())
    cal = np.array([0.62, -1.3e-3, 2.1e-6, 2.4e-4, 3.0e-7, -1.1e-9], np.float32)
    (d / "paint.cal").write_text("".join("%s = %.9g\n" % (k, v) for k, v in zip("abcdef", cal)))
    (d / "out").mkdir()
    (d / "job").mkdir()
    (d / "deck.inp").write_text(
        f"@general\n\ttest = t\n\trun = 1\n\tsequence = 1\n\ttunnel = ames_unitary\n@all\n\tgrid = {d}/model.tri\n\tsds = {d}/run.wtd\n"
        f"\ttargets = {d}/model.tgts\n{cams_txt}"
        f"@options\n\ttarget_patcher = none\n\tregistration = none\n\tfilter = none\n\tfilter_size = 1\n\toblique_angle = 70\n"
        f"\toverlap = average_view\n\tnumber_frames = {F}\n@output\n\tdir = {d}/out\n\tname = r\n")
    r = subprocess.run([up.build.build_setup_tool(), "-input_file", str(d / "deck.inp"), "-paint_cal", str(d / "paint.cal"),
                        "-job_dir", str(d / "job")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([up.build.build_host(), "-job_dir", str(d / "job"), "-out_dir", str(d / "out"), "-chunk", "8"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr

    subprocess.run([up.build.build_grid_probe(), str(d / "model.tri"), str(d / "g")], check=True, capture_output=True)
    nrm = np.fromfile(d / "g.nrm", np.float32).reshape(-1, 3)
    nrm_w = np.fromfile(d / "g.nrmw", np.float32).reshape(-1, 3)       # Node::get_normal(): what the camera weights take (projection.ipp:990)
    xyz = sc["xyz"].astype(np.float32)
    N = len(xyz)
    thresh = float(np.float32((180.0 - 70.0) * np.pi / 180.0))
    codes, angs = [], []
    for c, rv in enumerate(rvecs):
        pc = subprocess.run([up.build.build_setup_tool(), "-cal", str(d / f"cam{c + 1:02d}.json"), "-print_cal"], capture_output=True, text=True)
        parsed = {l.split()[0]: np.array(l.split()[1:], float) for l in pc.stdout.splitlines()}
        ocam = orc.make_camera(parsed["rvec"], parsed["tvec"], sc["K"], sc["dist"], W, H)
        codes.append(orc.create_projection(ocam, xyz, nrm, np.ones(N, np.uint8), sc["tri"], thresh)[0])
        angs.append(_angles(xyz, nrm_w, orc.cam_center(ocam)))
    seen = np.stack([c >= 0 for c in codes])
    both = seen[0] & seen[1]
    assert both.sum() > 30 and (seen[0] ^ seen[1]).sum() > 30
    total = (angs[0] + angs[1]).astype(np.float32)
    csr = []
    for c in range(2):
        rowptr, col, val = orc.projection_csr(codes[c])
        w = np.where(both, (angs[c] / total).astype(np.float32), np.float32(1))[seen[c]]
        csr.append((rowptr, col, (val * w).astype(np.float32)))
        assert np.array_equal(np.fromfile(d / "job" / f"cam{c}.col", np.int32), col)
        assert np.array_equal(np.fromfile(d / "job" / f"cam{c}.val", np.float32).view(np.uint32), csr[c][2].view(np.uint32))
    case = Case.__new__(Case)
    case.C, case.N, case.F, case.H, case.W = 2, N, F, H, W
    case.interp, case.degree, case.fmt, case.filter_kind, case.filter_size = 1, 6, "p12", 0, 0
    case.frames, case.csr, case.warp, case.patch_lists, case.overlap, case.synth = frames, csr, None, None, None, synth
    case.cal, case.qbar, case.ps = cal, np.float32(657.9153), np.float32(1332.0421)
    case.steady, case.temp = np.zeros(N, np.float32), np.full(N, 88.125, np.float32)
    ref = run_oracle(orc, case)
    got = np.fromfile(d / "out" / "intensity_transpose", np.float32).reshape(N, F)
    assert same_bits(got, ref["itrans"])
    assert same_bits(np.fromfile(d / "out" / "coverage", np.float32), ref["coverage"])
    assert same_bits(np.fromfile(d / "out" / "gain", np.float32), ref["gain"])


def sphere_zones(radius=5.0, J=49, K=13):
    """a sphere band about the x axis as two structured zones: the zones share the "equator" row (the circle x = 0,
    which passes in front of the camera at z = -radius) and each zone closes on itself in longitude (j = 0 and j = J-1
    coincide: a wrapped zone, its seam also facing the camera); the face normals e_j x e_k point outwards"""
    lon = np.linspace(0.0, 2 * np.pi, J)
    zones = []
    for lat0, lat1 in ((-1.2, 0.0), (0.0, 1.2)):
        lat = np.linspace(lat0, lat1, K)
        la, lo = np.meshgrid(lat, lon, indexing="ij")                       # [K, J]
        p = np.stack([radius * np.sin(la), radius * np.cos(la) * np.sin(lo), -radius * np.cos(la) * np.cos(lo)], -1).astype(np.float32)
        p[:, J - 1] = p[:, 0]
        zones.append((J, K, p.reshape(-1, 3)))
    zones[1][2][:J] = zones[0][2][-J:]                                      # shared equator row, bit-identical
    return zones


@pytest.mark.gpu
def test_deck_structured_grid_with_seams(up, orc, gpu, tmp_path):
    """SURVEY 8d config-4 style: a multi-zone plot3d model whose seam nodes overlap (zone-to-zone and wrapped zones).
    Superceded nodes get no projection row; every frame's solution is copied onto them from the lowest node of their
    group (P3DModel::adjust_solution -> remap.i32 -> upsp_gpu_set_overlap_remap)."""
    import cv2
    from oracle import p3d_overlap
    from test_p3d_model import write_p3d
    synth = up.synth
    sc = synth.make_projection_scene(n_lat=8, n_lon=16, seed=11)            # camera only
    W, H, F = sc["width"], sc["height"], 8
    d = tmp_path
    zones = sphere_zones()
    write_p3d(d / "model.grid", zones)
    xyz = np.concatenate([z[2] for z in zones]).astype(np.float32)
    sizes = [(z[0], z[1]) for z in zones]
    N = len(xyz)
    _cal_json(d / "cam01.json", cv2.Rodrigues(np.asarray(sc["rvec"], float))[0], sc["tvec"], sc["K"], sc["dist"], (W, H))
    frames = synth.make_frames(F, H, W, seed=9)[0]
    synth.pack_12bit(frames.reshape(F, -1)).tofile(d / "v1.mraw")
    (d / "v1.cih").write_text("#Camera Information Header\r\nRecord Rate(fps) : 1000\r\nTotal Frame : %d\r\n"
                              "Image Width : %d\r\nImage Height : %d\r\nColor Bit : 12\r\n" % (F, W, H))
    (d / "run.wtd").write_text(open(os.path.join(GOLDEN, "sample.wtd")).read())
    (d / "model.tgts").write_text(open(os.path.join(GOLDEN, "sample.tgts")).read())
    cal = np.array([0.62, -1.3e-3, 2.1e-6, 2.4e-4, 3.0e-7, -1.1e-9], np.float32)
    (d / "paint.cal").write_text("".join("%s = %.9g\n" % (k, v) for k, v in zip("abcdef", cal)))
    (d / "out").mkdir()
    (d / "job").mkdir()
    (d / "deck.inp").write_text(
        f"@general\n\ttest = t\n\trun = 1\n\tsequence = 1\n\ttunnel = ames_unitary\n@all\n\tgrid = {d}/model.grid\n\tsds = {d}/run.wtd\n"
        f"\ttargets = {d}/model.tgts\n@camera\n\tnumber = 1\n\tfilename = {d}/v1.mraw\n\tcalibration = {d}/cam01.json\n"
        f"@options\n\ttarget_patcher = none\n\tregistration = none\n\tfilter = none\n\tfilter_size = 1\n\toblique_angle = 70\n"
        f"\tnumber_frames = {F}\n@output\n\tdir = {d}/out\n\tname = r\n")
    r = subprocess.run([up.build.build_setup_tool(), "-input_file", str(d / "deck.inp"), "-paint_cal", str(d / "paint.cal"),
                        "-job_dir", str(d / "job")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([up.build.build_host(), "-job_dir", str(d / "job"), "-out_dir", str(d / "out"), "-chunk", "8"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr

    overlap, _, _ = p3d_overlap.identify_overlap(xyz, sizes, 1e-3)
    remap = orc.overlap_remap(N, overlap)
    assert np.array_equal(np.fromfile(d / "job" / "remap.i32", np.int32), remap) and (remap != np.arange(N)).sum() >= 49 + 2 * 12
    subprocess.run([up.build.build_inputs_probe(), "overlap", str(d / "model.grid"), "0.001", str(d / "g")], check=True, capture_output=True)
    nrm = np.fromfile(d / "g.nrm", np.float32).reshape(-1, 3)
    tri = np.fromfile(d / "g.trinodes", np.int32).reshape(-1, 3)
    assert np.array_equal(tri.ravel(), p3d_overlap.extract_tri_nodes(sizes))
    pc = subprocess.run([up.build.build_setup_tool(), "-cal", str(d / "cam01.json"), "-print_cal"], capture_output=True, text=True)
    parsed = {l.split()[0]: np.array(l.split()[1:], float) for l in pc.stdout.splitlines()}
    ocam = orc.make_camera(parsed["rvec"], parsed["tvec"], sc["K"], sc["dist"], W, H)
    is_data = (remap == np.arange(N)).astype(np.uint8)
    code, _ = orc.create_projection(ocam, xyz, nrm, is_data, tri, float(np.float32((180.0 - 70.0) * np.pi / 180.0)))
    rowptr, col, val = orc.projection_csr(code)
    assert (code >= 0).sum() > 100 and not np.any(code[is_data == 0] >= 0)
    assert np.array_equal(np.fromfile(d / "job" / "cam0.rowptr", np.int32), rowptr)
    assert np.array_equal(np.fromfile(d / "job" / "cam0.col", np.int32), col)
    case = Case.__new__(Case)
    case.C, case.N, case.F, case.H, case.W = 1, N, F, H, W
    case.interp, case.degree, case.fmt, case.filter_kind, case.filter_size = 1, 6, "p12", 0, 0
    case.frames, case.csr, case.warp, case.patch_lists, case.overlap, case.synth = [frames], [(rowptr, col, val)], None, None, overlap, synth
    case.cal, case.qbar, case.ps = cal, np.float32(657.9153), np.float32(1332.0421)
    case.steady, case.temp = np.zeros(N, np.float32), np.full(N, 88.125, np.float32)
    ref = run_oracle(orc, case)
    got = np.fromfile(d / "out" / "intensity_transpose", np.float32).reshape(N, F)
    assert same_bits(got, ref["itrans"])
    seam = np.nonzero((remap != np.arange(N)) & np.isfinite(got[:, 0]))[0]
    assert len(seam) > 5 and np.array_equal(got[seam], got[remap[seam]])      # seam nodes carry their group's values
    assert same_bits(np.fromfile(d / "out" / "intensity_avg", np.float32), ref["avg"])
    assert same_bits(np.fromfile(d / "out" / "coverage", np.float32), ref["coverage"])
