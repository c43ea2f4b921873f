"""Phase-0 product: the pixel-to-node projection matrix.
CPU: the oracle's camera model against cv2.projectPoints / cv2.Rodrigues golden vectors (the call
of CameraCal::map_points_to_image), and properties of the visibility test on a synthetic scene.
GPU: upsp_op_project_points / upsp_op_create_projection against the oracle, bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "setup_golden.npz"))


def _cam_args(row):
    K = np.array([[row[6], 0, row[8]], [0, row[7], row[9]], [0, 0, 1.0]])
    return row[0:3], row[3:6], K, row[10:18]


def test_oracle_camera_model_matches_cv2(orc, gold):
    for cam_row, pts, uv, rot in zip(gold["cams"], gold["pts"], gold["uv"], gold["rot"]):
        rvec, tvec, K, dist = _cam_args(cam_row)
        cam = orc.make_camera(rvec, tvec, K, dist, 1024, 1024)
        got = orc.project_points(cam, pts)
        assert np.array_equal(got.view(np.uint32), uv.view(np.uint32))       # bit-exact vs cv2.projectPoints
        R = np.empty(9)
        orc.lib().orc_rodrigues((C.c_double * 3)(*rvec), R.ctypes.data_as(C.c_void_p))
        assert np.array_equal(R, rot)                                          # bit-exact vs cv2.Rodrigues
        centre = -(rot.reshape(3, 3).T @ tvec)                                 # CameraCal::get_cam_center
        assert np.allclose(orc.cam_center(cam), centre, rtol=0, atol=1e-5)


def test_oracle_projection_properties(orc):
    import upsp_b200
    sc = upsp_b200.synth.make_projection_scene(n_lat=20, n_lon=40)
    cam = orc.make_camera(sc["rvec"], sc["tvec"], sc["K"], sc["dist"], sc["width"], sc["height"])
    code, uv = orc.create_projection(cam, sc["xyz"], sc["normals"], sc["is_data"], sc["tri"], sc["thresh"])
    acc = code >= 0
    assert 50 < acc.sum() < len(code) // 2                    # only the side facing the camera
    assert not np.any(acc & (sc["is_data"] == 0))             # non-data nodes never get an entry
    pts = orc.project_points(cam, sc["xyz"])
    assert np.array_equal(code[acc], (np.floor(pts[acc, 1] + 0.5) * sc["width"] + np.floor(pts[acc, 0] + 0.5)).astype(np.int32))
    assert np.array_equal(uv[acc], pts[acc] / np.float32([sc["width"], sc["height"]]))
    assert np.all(uv[~acc] == 0)
    # nodes facing away from the camera are rejected by the ray cast (they are hit from behind)
    centre = orc.cam_center(cam)
    facing = np.einsum("ij,ij->i", sc["normals"], sc["xyz"] - centre[None, :]) < 0
    assert not np.any(acc & ~facing)
    # removing the occluder's triangles makes more of the model visible, never less
    n_occ_nodes = 8 * 16 + 2
    n_occ_tris = 2 * 16 + 2 * 16 * 7
    code2, _ = orc.create_projection(cam, sc["xyz"], sc["normals"], sc["is_data"], sc["tri"][:-n_occ_tris], sc["thresh"])
    main = slice(0, len(code) - n_occ_nodes)
    assert np.all((code2[main] >= 0) | ~(code[main] >= 0)) and (code2[main] >= 0).sum() > (code[main] >= 0).sum()
    rowptr, col, val = orc.projection_csr(code)
    assert rowptr[-1] == acc.sum() == len(col) and np.all(val == 1.0)


@pytest.mark.gpu
def test_gpu_project_points_matches_cv2(up, gpu, gold):
    for cam_row, pts, uv in zip(gold["cams"], gold["pts"], gold["uv"]):
        rvec, tvec, K, dist = _cam_args(cam_row)
        got = up.op_project_points(up.camera_model(rvec, tvec, K, dist, 1024, 1024), pts)
        assert np.array_equal(got.view(np.uint32), uv.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("n_lat,n_lon,behind", [(12, 24, False), (40, 80, False), (30, 60, True)])
def test_gpu_create_projection_matches_oracle(up, orc, gpu, n_lat, n_lon, behind):
    """Pixel-to-node indices and sparsity bit-exact (BASELINE north star), (u, v) bit-exact.  `behind`
    moves the camera so close that part of the model lies behind the camera plane: those triangles
    are on the 'tested by every ray' list and rays pointing backwards take the brute-force path."""
    import upsp_b200
    sc = upsp_b200.synth.make_projection_scene(n_lat=n_lat, n_lon=n_lon, seed=n_lat)
    if behind:
        sc["tvec"] = np.array([6.0, 0.0, 6.0])            # camera beside the model: its plane cuts through it
        sc["K"] = np.array([[300.0, 0, 255.3], [0, 300.0, 250.8], [0, 0, 1]])
    ocam = orc.make_camera(sc["rvec"], sc["tvec"], sc["K"], sc["dist"], sc["width"], sc["height"])
    gcam = up.camera_model(sc["rvec"], sc["tvec"], sc["K"], sc["dist"], sc["width"], sc["height"])
    ref_code, ref_uv = orc.create_projection(ocam, sc["xyz"], sc["normals"], sc["is_data"], sc["tri"], sc["thresh"])
    code, uv = up.op_create_projection(gcam, sc["xyz"], sc["normals"], sc["is_data"], sc["tri"], sc["thresh"])
    assert (ref_code >= 0).sum() > 20
    assert np.array_equal(code, ref_code)
    assert np.array_equal(uv.view(np.uint32), ref_uv.view(np.uint32))
    with pytest.raises(up.UpspGpuError):
        bad = sc["tri"].copy()
        bad[0, 0] = len(sc["xyz"])
        up.op_create_projection(gcam, sc["xyz"], sc["normals"], sc["is_data"], bad, sc["thresh"])
