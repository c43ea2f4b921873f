"""psp_setup_b200 (host/psp_setup_b200.cpp, camera_cal.hpp): phase 0 for one camera and a .tri grid.
CPU: the calibration reader (read_json_camera_calibration, cpp/lib/CameraCal.cpp:18-54) and its
cv::Rodrigues(matrix -> vector) against cv2 golden vectors.  GPU: grid + calibration -> projection
matrix files equal the oracle's create_projection_mat on the same inputs."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_grid_readers import write_tri


def _cal_json(path, R, tvec, K, dist, size):
    json.dump(dict(cameraMatrix=np.asarray(K).tolist(), distCoeffs=list(map(float, dist)), rmat=np.asarray(R).tolist(),
                   tvec=list(map(float, tvec)), imageSize=list(size)), open(path, "w"))


def _print_cal(exe, path):
    r = subprocess.run([exe, "-cal", str(path), "-print_cal"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return {l.split()[0]: np.array(l.split()[1:], float) for l in r.stdout.splitlines()}


def test_calibration_reader_matches_cv2(up, tmp_path):
    exe = up.build.build_setup_tool()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "setup_golden.npz"))
    for row, rot in zip(gold["cams"], gold["rot"]):
        K = [[row[6], 0, row[8]], [0, row[7], row[9]], [0, 0, 1]]
        nd = 8 if np.any(row[15:18]) else 5 if row[14] else 4
        _cal_json(tmp_path / "c.json", rot.reshape(3, 3), row[3:6], K, row[10:10 + nd], (640, 480))
        got = _print_cal(exe, tmp_path / "c.json")
        assert np.allclose(got["rvec"], row[0:3], rtol=0, atol=1e-12)      # cv2.Rodrigues(R) == the rvec R came from
        assert np.array_equal(got["tvec"], row[3:6]) and np.array_equal(got["K"], [row[6], row[7], row[8], row[9]])
        assert np.array_equal(got["dist"][:nd], row[10:10 + nd]) and not np.any(got["dist"][nd:])
        assert np.array_equal(got["imageSize"], [640, 480])
    # identity and half-turn rotations (the special cases of cv::Rodrigues)
    for R, want in ((np.eye(3), [0, 0, 0]), (np.diag([1.0, -1.0, -1.0]), [np.pi, 0, 0]), (np.diag([-1.0, -1.0, 1.0]), [0, 0, np.pi])):
        _cal_json(tmp_path / "c.json", R, [0, 0, 1], np.eye(3), [0, 0, 0, 0], (8, 8))
        assert np.allclose(_print_cal(exe, tmp_path / "c.json")["rvec"], want, atol=1e-12)


def test_setup_tool_fails_loudly(up, tmp_path):
    exe = up.build.build_setup_tool()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
    r = subprocess.run([exe, "-cal", str(tmp_path / "none.json"), "-print_cal"], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot open camera calibration file" in r.stderr
    (tmp_path / "bad.json").write_text('{"cameraMatrix": [[1,0,0],[0,1,0],[0,0,1]], "distCoeffs": [0,0,0], "rmat": [[1,0,0],[0,1,0],[0,0,1]], '
                                       '"tvec": [0,0,1], "imageSize": [4,4]}')
    r = subprocess.run([exe, "-cal", str(tmp_path / "bad.json"), "-print_cal"], capture_output=True, text=True)
    assert r.returncode == 1 and "4,5,or 8 coefficients" in r.stderr


@pytest.mark.gpu
def test_setup_tool_matches_oracle(up, orc, gpu, tmp_path):
    import cv2
    exe = up.build.build_setup_tool()
    sc = up.synth.make_projection_scene(n_lat=24, n_lon=48, seed=11)
    write_tri(tmp_path / "model.tri", sc["xyz"], sc["tri"], np.ones(len(sc["tri"]), np.int32))
    R = cv2.Rodrigues(np.asarray(sc["rvec"], float))[0]
    _cal_json(tmp_path / "cam.json", R, sc["tvec"], sc["K"], sc["dist"], (sc["width"], sc["height"]))
    r = subprocess.run([exe, "-grid", str(tmp_path / "model.tri"), "-cal", str(tmp_path / "cam.json"), "-out_dir", str(tmp_path),
                        "-oblique_angle", "70"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cal = {l.split()[0]: np.array(l.split()[1:], float) for l in r.stdout.splitlines() if l.split()[0] in ("rvec", "tvec")}
    # the oracle run: same grid arrays, normals as calcNormals defines them, the calibration the tool parsed
    subprocess.run([up.build.build_grid_probe(), str(tmp_path / "model.tri"), str(tmp_path / "d")], check=True, capture_output=True)
    nrm = np.fromfile(tmp_path / "d.nrm", np.float32).reshape(-1, 3)
    ocam = orc.make_camera(cal["rvec"], cal["tvec"], sc["K"], sc["dist"], sc["width"], sc["height"])
    thresh = np.float32((180.0 - 70.0) * np.pi / 180.0)
    code, uv = orc.create_projection(ocam, sc["xyz"], nrm, np.ones(len(sc["xyz"]), np.uint8), sc["tri"], float(thresh))
    rowptr, col, val = orc.projection_csr(code)
    assert (code >= 0).sum() > 100
    assert np.array_equal(np.fromfile(tmp_path / "cam0.rowptr", np.int32), rowptr)
    assert np.array_equal(np.fromfile(tmp_path / "cam0.col", np.int32), col)
    assert np.array_equal(np.fromfile(tmp_path / "cam0.val", np.float32), val)
    assert np.array_equal(np.fromfile(tmp_path / "cam01-uv", np.float32).view(np.uint32), uv.ravel().view(np.uint32))
    assert f"accepted {len(col)}" in r.stdout
