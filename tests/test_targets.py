"""host/targets.hpp: which fiducial targets a camera sees, where, and how large (getTargets
cpp/exec/psp_process.cpp:56-114, get_target_diameters :116-165), as run by `psp_setup_b200 -input_file` when the deck
asks for the polynomial patcher.  Image positions are held bit-exact against cv2.projectPoints (the OpenCV entry point
behind CameraCal::map_point_to_image); visibility and diameters against the float64 model oracle/targets.py.  CPU only."""
import subprocess

import numpy as np
import pytest

from oracle import targets as otargets
from test_grid_readers import write_tri
from test_setup_deck import make_inputs, setup
from test_setup_tool import _cal_json


def test_visible_targets_from_deck(up, tmp_path):
    cv2 = pytest.importorskip("cv2")
    sc = up.synth.make_projection_scene(n_lat=24, n_lon=48, seed=11)
    make_inputs(up, tmp_path, "tri")
    write_tri(tmp_path / "model.tri", sc["xyz"], sc["tri"], np.ones(len(sc["tri"]), np.int32))
    W, H = 32, 16                                              # the committed camera file; K scaled to that frame
    K = np.array([[150.0, 0, 15.6], [0, 149.0, 7.9], [0, 0, 1]])
    _cal_json(tmp_path / "cam01.json", cv2.Rodrigues(np.asarray(sc["rvec"], float))[0], sc["tvec"], K, sc["dist"], (W, H))
    # targets on the big sphere (radius 5 about the origin; the camera sits near z = -30): facing the camera, behind
    # the small occluding sphere, grazing, on the far side, outside the 32 x 16 frame
    rng = np.random.default_rng(2)
    dirs = [(0.02, 0.03, -1), (0.15, -0.05, -1), (-0.2, 0.1, -1), (0.30, 0.20, -1), (-0.05, 0.12, -1), (0.22, 0.14, -1),
            (0.0, 0.0, 1), (0.3, 0.1, 1), (0.9, 0.0, -0.45), (0.0, 0.6, -0.8), (-0.4, -0.1, -1), (0.05, -0.18, -1)]
    lines, targets = [], []
    for i, dvec in enumerate(dirs):
        dvec = np.asarray(dvec, float) + rng.normal(0, 0.01, 3)
        p = 5.0 * dvec / np.linalg.norm(dvec)
        diam = 0.4 + 0.05 * i
        lines.append("%4d %10.4f %10.4f %10.4f 0.0 0.0 1.0 %6.3f 1 2 3 st%02d\n" % (i + 1, p[0], p[1], p[2], diam, i + 1))
        targets.append(tuple(np.float32(float(s)) for s in lines[-1].split()[1:4]) + (np.float32(float(lines[-1].split()[7])),))
    (tmp_path / "model.tgts").write_text("#hdr\n*Targets\n" + "".join(lines[:9]) + "*Fiducials\n" + "".join(lines[9:]) + "*Taps\n 1 0 0 0 0 0 1 0.1 1 1 1 t\n")
    (tmp_path / "deck.inp").write_text((tmp_path / "deck.inp").read_text().replace("target_patcher = none", "target_patcher = polynomial"))
    r, job = setup(up, tmp_path, "-no_projection", "-target_diam_sf", "1.2")
    got = np.loadtxt(job / "cam0.targets", dtype=np.float32).reshape(-1, 3)

    subprocess.run([up.build.build_grid_probe(), str(tmp_path / "model.tri"), str(tmp_path / "g")], check=True, capture_output=True)
    nrm = np.fromfile(tmp_path / "g.nrm", np.float32).reshape(-1, 3)          # Model::get_n(): visibility of getTargets
    nrm_w = np.fromfile(tmp_path / "g.nrmw", np.float32).reshape(-1, 3)       # Node::get_normal(): get_target_diameters
    pc = subprocess.run([up.build.build_setup_tool(), "-cal", str(tmp_path / "cam01.json"), "-print_cal"], capture_output=True, text=True)
    rvec = np.array(pc.stdout.splitlines()[0].split()[1:], float)
    want = otargets.visible_targets(cv2, targets, sc["xyz"], nrm, sc["tri"], rvec, sc["tvec"], K, sc["dist"], W, H, 70.0, 1.2, node_normals=nrm_w)
    assert 3 <= len(want) < len(dirs) - 3                      # some of each kind were rejected
    assert all(w[4] > 0.02 for w in want), "a target sits on a decision boundary; move it"
    assert f"{len(want)} visible targets" in r.stdout and len(got) == len(want)
    for g, w in zip(got, want):
        assert g[0] == w[1] and g[1] == w[2]                   # == cv2.projectPoints, bit for bit
        assert abs(float(g[2]) - w[3]) <= 2e-5 * w[3]
