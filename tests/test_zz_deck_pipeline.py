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
    r = subprocess.run([up.build.build_setup_tool(), "-input_file", str(d / "deck.inp"), "-paint_cal", str(d / "paint.cal"),
                        "-job_dir", str(d / "job")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr

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
    assert np.array_equal(np.fromfile(d / "job" / "cam0.rowptr", np.int32), rowptr)
    assert np.array_equal(np.fromfile(d / "job" / "cam0.col", np.int32), col)
    assert np.array_equal(np.fromfile(d / "job" / "cam0.val", np.float32), val)          # one camera: weights stay 1

    r = subprocess.run([up.build.build_host(), "-job_dir", str(d / "job"), "-out_dir", str(d / "out"), "-chunk", "8"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr

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
