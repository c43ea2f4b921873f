"""psp_setup_b200 deck mode (host/psp_setup_b200.cpp: run_deck): psp_process's own command line and input deck
(ParseOpts cpp/exec/psp_process.cpp:1192-1310, InitializeVideoStreams :392-470, phase-2 start-up :2270-2385) turned
into a job directory of psp_process_b200.  CPU: everything except the projection matrix (`-no_projection`)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN
from test_grid_readers import write_tri
from test_p3d_model import reference_fixture, write_p3d
from test_setup_tool import _cal_json


def job_kv(path):
    out = {}
    for line in open(path):
        if "=" in line and not line.startswith("#"):
            k, _, v = line.partition("=")
            out[k.strip()] = v.strip()
    return out


def make_inputs(up, d, grid="tri", registration="none", frames=-1, cameras=1, extra_all=""):
    """a run directory in the reference's layout; the camera file is the committed 2-frame 32x16 Photron pair"""
    if grid == "tri":
        xyz, _, tri = up.synth.make_sphere_mesh(8, 16, 2.0, (0, 0, 6.0), seed=3)
        write_tri(d / "model.tri", xyz.astype(np.float32), tri, np.ones(len(tri), np.int32))
        gname, n = "model.tri", len(xyz)
    else:
        write_p3d(d / "model.grid", reference_fixture())
        gname, n = "model.grid", 52
    (d / "run.wtd").write_text(open(os.path.join(GOLDEN, "sample.wtd")).read())
    (d / "model.tgts").write_text(open(os.path.join(GOLDEN, "sample.tgts")).read())
    (d / "paint.cal").write_text("a = 1.5\nb = -0.002\nc = 1e-6\nd = 0.5\ne = 0.001\nf = -2e-7\n")
    for c in range(cameras):
        _cal_json(d / f"cam{c + 1:02d}.json", np.eye(3), [0.1 * c, 0, 0], [[40, 0, 16], [0, 40, 8], [0, 0, 1]], [0, 0, 0, 0], (32, 16))
        for ext in ("mraw", "cih"):
            (d / f"v{c + 1}.{ext}").write_bytes(open(os.path.join(GOLDEN, "tiny12." + ext), "rb").read())
    (d / "out").mkdir(exist_ok=True)
    cams = "".join(f"@camera\n\tnumber = {c + 1}\n\tfilename = $d/v{c + 1}.mraw\n\tcalibration = $d/cam{c + 1:02d}.json\n" for c in range(cameras))
    deck = (f"%Version 0.0\n@general\n\ttest = t-b200\n\trun = 12\n\tsequence = 3\n\ttunnel = ames_unitary\n@vars\n\td = {d}\n"
            f"@all\n\tgrid = $d/{gname}\n\tsds = $d/run.wtd\n\ttargets = $d/model.tgts\n{extra_all}{cams}"
            f"@options\n\ttarget_patcher = none\n\tregistration = {registration}\n\tfilter = none\n\tfilter_size = 1\n"
            f"\toblique_angle = 70\n\tnumber_frames = {frames}\n\toverlap = best_view\n@output\n\tdir = $d/out\n\tname = run12\n")
    (d / "deck.inp").write_text(deck)
    return n


def setup(up, d, *extra, ok=True):
    job = d / "job"
    job.mkdir(exist_ok=True)
    r = subprocess.run([up.build.build_setup_tool(), "-input_file", str(d / "deck.inp"), "-paint_cal", str(d / "paint.cal"),
                        "-job_dir", str(job)] + [str(a) for a in extra], capture_output=True, text=True)
    assert (r.returncode == 0) == ok, r.stdout + r.stderr
    return r, job


def test_deck_to_job_unstructured(up, tmp_path):
    n = make_inputs(up, tmp_path, "tri", registration="pixel")
    r, job = setup(up, tmp_path, "-no_projection")
    kv = job_kv(job / "job.txt")
    assert kv["cameras"] == "1" and kv["width"] == "32" and kv["height"] == "16" and kv["msize"] == str(n)
    assert kv["number_frames"] == "2"                                   # -1 in the deck: every frame all cameras have
    assert kv["video0"] == str(tmp_path / "v1.mraw") and kv["registration"] == "pixel" and kv["target_patcher"] == "none"
    assert kv["bound_thickness"] == "2" and kv["buffer_thickness"] == "1" and kv["degree"] == "6"
    assert np.float32(kv["qbar"]) == np.float32(657.9153) and np.float32(kv["ps"]) == np.float32(1332.0421)
    assert [np.float32(kv["cal_" + k]) for k in "abcdef"] == [np.float32(v) for v in (1.5, -0.002, 1e-6, 0.5, 0.001, -2e-7)]
    assert np.array_equal(np.fromfile(job / "steady.f32", np.float32), np.zeros(n, np.float32))       # wind-off
    assert np.array_equal(np.fromfile(job / "model_temp.f32", np.float32), np.full(n, 88.125, np.float32))   # TCAVG
    assert not (job / "remap.i32").exists() and np.fromfile(job / "xyz.f32", np.float32).size == 3 * n
    assert "Will process (2) frames" in r.stdout and "thermocouple average" in r.stdout


def test_deck_to_job_structured(up, tmp_path):
    n = make_inputs(up, tmp_path, "p3d", frames=1)
    vals = (np.arange(n, dtype=np.float32) * 0.01 - 0.2).astype(np.float32)
    with open(tmp_path / "steady.f", "wb") as f:                        # no record markers
        f.write(struct.pack("<i", 3) + b"".join(struct.pack("<4i", j, k, 1, 1) for j, k in ((4, 5), (3, 4), (5, 4))) + vals.tobytes())
    r, job = setup(up, tmp_path, "-no_projection", "-steady_p3d", tmp_path / "steady.f")
    kv = job_kv(job / "job.txt")
    assert kv["msize"] == "52" and kv["number_frames"] == "1"
    remap = np.fromfile(job / "remap.i32", np.int32)
    assert sorted(np.nonzero(remap != np.arange(52))[0]) == [20, 23, 26, 29, 41, 46, 51]
    assert np.array_equal(np.fromfile(job / "steady.f32", np.float32), vals)
    (tmp_path / "short.f").write_bytes(struct.pack("<i", 1) + struct.pack("<4i", 2, 2, 1, 1) + np.zeros(4, np.float32).tobytes())
    r, _ = setup(up, tmp_path, "-no_projection", "-steady_p3d", tmp_path / "short.f", ok=False)
    assert "Steady-state function file inconsistent with grid (expect 52 values, got 4)" in r.stderr


def test_deck_mode_rejections(up, tmp_path):
    make_inputs(up, tmp_path, "tri", frames=3)
    r, _ = setup(up, tmp_path, "-no_projection", ok=False)
    assert "(3) frames requested but only (2) frames available" in r.stderr
    r, _ = setup(up, tmp_path, "-no_projection", "-frames", "2")          # -frames overrides the deck
    make_inputs(up, tmp_path, "tri", registration="point")
    assert "Unsupported registration type" in setup(up, tmp_path, "-no_projection", ok=False)[0].stderr
    make_inputs(up, tmp_path, "tri")
    (tmp_path / "deck.inp").write_text((tmp_path / "deck.inp").read_text().replace("ames_unitary", "langley"))
    assert "Unrecognized tunnel name 'langley'" in setup(up, tmp_path, "-no_projection", ok=False)[0].stderr
    make_inputs(up, tmp_path, "tri")
    (tmp_path / "deck.inp").write_text((tmp_path / "deck.inp").read_text().replace("filter_size = 1", "filter_size = 4"))
    assert "Filter size must be odd" in setup(up, tmp_path, "-no_projection", ok=False)[0].stderr
    make_inputs(up, tmp_path, "tri")
    os.remove(tmp_path / "run.wtd")
    assert "SDS file" in setup(up, tmp_path, "-no_projection", ok=False)[0].stderr
    make_inputs(up, tmp_path, "tri")
    r = subprocess.run([up.build.build_setup_tool(), "-input_file", str(tmp_path / "deck.inp"), "-job_dir", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "Must specify -paint_cal" in r.stderr
    make_inputs(up, tmp_path, "tri")
    (tmp_path / "steady.f").write_bytes(b"\0" * 64)
    assert "needs -steady_grid" in setup(up, tmp_path, "-no_projection", "-steady_p3d", tmp_path / "steady.f", ok=False)[0].stderr


def test_deck_wind_on_unstructured(up, tmp_path):
    """-steady_p3d + -steady_grid with a .tri model: Cp interpolated from the structured steady grid (k = 10, 1/d^2)"""
    from test_interpolation import numpy_interpolate
    n = make_inputs(up, tmp_path, "tri", frames=1)
    J, K = 9, 7
    th, z = np.meshgrid(np.linspace(0, np.pi, J), np.linspace(4.5, 7.5, K))
    sg = np.stack([2.1 * np.cos(th), 2.1 * np.sin(th), z], -1).reshape(-1, 3).astype(np.float32)
    write_p3d(tmp_path / "steady.grid", [(J, K, sg)])
    vals = np.sin(np.arange(J * K, dtype=np.float32)).astype(np.float32)
    with open(tmp_path / "steady.f", "wb") as f:
        f.write(struct.pack("<i", 1) + struct.pack("<4i", J, K, 1, 1) + vals.tobytes())
    r, job = setup(up, tmp_path, "-no_projection", "-steady_p3d", tmp_path / "steady.f", "-steady_grid", tmp_path / "steady.grid")
    xyz = np.fromfile(job / "xyz.f32", np.float32).reshape(-1, 3)
    want = numpy_interpolate(sg, np.ones(J * K, bool), vals, xyz, 10)
    got = np.fromfile(job / "steady.f32", np.float32)
    assert len(got) == n and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_deck_normals_override_and_active_components(up, tmp_path):
    # structured: normals csv overrides node normals (file_readers.ipp:12-90), inactive zone -> non-data nodes
    make_inputs(up, tmp_path, "p3d", frames=1, extra_all="\tnormals = $d/n.csv\n\tactive_comps = $d/comps.csv\n")
    (tmp_path / "n.csv").write_text("nidx, x_norm, y_norm, z_norm\n3, 0.6, 0.0, 0.8\n40, -1, 0, 0\n")
    (tmp_path / "comps.csv").write_text("component,active\n0,1\n1,0\n")
    r, job = setup(up, tmp_path, "-no_projection")
    nrm = np.fromfile(job / "normals.f32", np.float32).reshape(-1, 3)
    assert np.array_equal(nrm[3], np.float32([0.6, 0.0, 0.8])) and np.array_equal(nrm[40], np.float32([-1, 0, 0]))
    assert np.array_equal(nrm[4], np.float32([0, 0, 1])) and "Overwrote 2/52 model surface normals" in r.stdout
    is_data = np.fromfile(job / "is_data.u8", np.uint8)
    want = np.ones(52, np.uint8)
    want[[20, 23, 26, 29, 41, 46, 51]] = 0                   # superceded nodes
    want[20:32] = 0                                          # zone 1 switched off
    assert np.array_equal(is_data, want)
    (tmp_path / "comps.csv").write_text("component,active\n0,1\n1,0\n2,1\n3,1\n")
    assert "Number of components in active component file" in setup(up, tmp_path, "-no_projection", ok=False)[0].stderr
    (tmp_path / "comps.csv").write_text("component,active\n0,1\n")
    (tmp_path / "n.csv").write_text("nidx, x_norm, y_norm\n3, 0.6, 0.0\n")
    assert "Could not parse z_norm" in setup(up, tmp_path, "-no_projection", ok=False)[0].stderr

    # unstructured: a node is switched off only when all its triangles carry the same (inactive) component;
    # the normals csv is refused with the reference's message and the run goes on
    xyz = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [2, 0, 0], [2, 1, 0]], np.float32) + np.float32([0, 0, 6])
    tri = np.array([[0, 1, 2], [0, 2, 3], [1, 4, 5], [1, 5, 2]], np.int32)
    make_inputs(up, tmp_path, "tri", frames=1, extra_all="\tnormals = $d/n.csv\n\tactive_comps = $d/comps.csv\n")
    write_tri(tmp_path / "model.tri", xyz, tri, np.array([1, 1, 2, 2], np.int32))
    (tmp_path / "comps.csv").write_text("component,active\n2,0\n")
    (tmp_path / "n.csv").write_text("nidx, x_norm, y_norm, z_norm\n0, 1, 0, 0\n")
    r, job = setup(up, tmp_path, "-no_projection", "-cutoff_x_max", "0.5")
    assert "can not specify normals CSV for a TriModel_" in r.stderr
    # nodes 4, 5 belong to component 2 only; 1, 2 touch both components; x > 0.5 removes 1, 2, 4, 5 anyway
    assert list(np.fromfile(job / "is_data.u8", np.uint8)) == [1, 0, 0, 1, 0, 0]
    r, job = setup(up, tmp_path, "-no_projection")
    assert list(np.fromfile(job / "is_data.u8", np.uint8)) == [1, 1, 1, 1, 0, 0]


def test_reference_command_line_forms(up, tmp_path):
    """cv::CommandLineParser style -key=value (what the reference's launcher writes, python/upsp/processing/tree.py:455-465)
    is accepted next to -key value; psp_process_b200's one-step mode checks the reference's required flags first."""
    make_inputs(up, tmp_path, "tri")
    job = tmp_path / "job"
    job.mkdir()
    r = subprocess.run([up.build.build_setup_tool(), f"-input_file={tmp_path / 'deck.inp'}", f"-paint_cal={tmp_path / 'paint.cal'}",
                        f"-job_dir={job}", "-frames=1", "-no_projection"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert job_kv(job / "job.txt")["number_frames"] == "1"
    exe = up.build.build_host()
    r = subprocess.run([exe, f"-input_file={tmp_path / 'deck.inp'}", f"-paint_cal={tmp_path / 'paint.cal'}"], capture_output=True, text=True)
    assert r.returncode == 1 and "Must specify -h5_out" in r.stderr
    r = subprocess.run([exe, f"-input_file={tmp_path / 'deck.inp'}", "-h5_out=x.h5"], capture_output=True, text=True)
    assert r.returncode == 1 and "Must specify -paint_cal" in r.stderr
    r = subprocess.run([exe, f"-input_file={tmp_path / 'nope.inp'}", "-h5_out=x.h5", "-paint_cal=p"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot be opened" in r.stderr


def test_deck_tri_grid_is_intersected(up, tmp_path):
    """psp_process loads the unstructured model with TriModel_(file, intersect = true) (psp_process.cpp:1384): duplicate nodes
    collapse before anything is sized, so msize and every per-node output follow the intersected numbering"""
    from test_grid_readers import numpy_intersect
    make_inputs(up, tmp_path, "tri", frames=1)
    xyz, _, tri = up.synth.make_sphere_mesh(8, 16, 2.0, (0, 0, 6.0), seed=3)
    xyz = xyz.astype(np.float32)
    dup = np.array([5, 17, 40])
    xyz2 = np.concatenate([xyz, xyz[dup]]).astype(np.float32)               # three nodes stored twice ...
    tri2 = tri.copy().astype(np.int32)
    for k, n in enumerate(dup):                                              # ... and used by one triangle each
        hit = np.argwhere(tri2 == n)[0]
        tri2[tuple(hit)] = len(xyz) + k
    comps = np.ones(len(tri2), np.int32)
    write_tri(tmp_path / "model.tri", xyz2, tri2, comps)
    r, job = setup(up, tmp_path, "-no_projection")
    wx, wt, _, _ = numpy_intersect(xyz2, tri2, comps)
    assert len(wx) == len(xyz) and "Found 3 non-unique points" in r.stdout
    assert job_kv(job / "job.txt")["msize"] == str(len(xyz))
    assert np.array_equal(np.fromfile(job / "xyz.f32", np.float32).reshape(-1, 3), wx)
