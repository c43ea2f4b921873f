"""host/projection_weights.hpp: multi-camera blending weights (adjust_projection_for_weights,
BestView / AverageViews, angle_between, identify_skipped_nodes; cpp/lib/projection.ipp:226-268,
857-880, 912-1078; cpp/utils/cv_extras.ipp:67-73) against a numpy restatement.  CPU only.
With two cameras per node the float sum of the angles does not depend on the order in which the
reference's priority queue pops the cameras, so the weights are compared bit for bit; with three
cameras the pop order among equal rows is the container's business and two ulp are allowed."""
import subprocess

import numpy as np
import pytest


def _angles(xyz, nrm, center):
    d = (xyz - np.float32(center)[None, :]).astype(np.float32)
    dot = (d[:, 0] * nrm[:, 0] + d[:, 1] * nrm[:, 1] + d[:, 2] * nrm[:, 2]).astype(np.float32)       # float dot product
    n1 = np.sqrt((d.astype(np.float64) ** 2).sum(1))
    n2 = np.sqrt((nrm.astype(np.float64) ** 2).sum(1))
    return np.arccos(dot.astype(np.float64) / n1 / n2).astype(np.float32)


def _scene(up, n_cams, seed):
    rng = np.random.default_rng(seed)
    xyz, nrm, _ = up.synth.make_sphere_mesh(14, 28, 4.0, seed=seed)
    n = len(xyz)
    centers = np.array([[30.0 * np.cos(a), 30.0 * np.sin(a), 5.0 * k] for k, a in enumerate(np.linspace(0, 1.2, n_cams))])
    cams = []
    for c in range(n_cams):
        seen = rng.random(n) < 0.6
        rowptr = np.concatenate([[0], np.cumsum(seen)]).astype(np.int32)
        cams.append((rowptr, rng.integers(0, 4096, int(seen.sum())).astype(np.int32), np.ones(int(seen.sum()), np.float32), seen))
    return xyz, nrm, centers, cams


def _run(up, tmp_path, xyz, nrm, centers, cams, mode):
    probe = up.build.build_weights_probe()
    xyz.astype("<f4").tofile(tmp_path / "xyz.f32")
    nrm.astype("<f4").tofile(tmp_path / "nrm.f32")
    centers.astype("<f8").tofile(tmp_path / "centers.f64")
    for c, (rowptr, col, val, _) in enumerate(cams):
        rowptr.tofile(tmp_path / f"cam{c}.rowptr")
        col.tofile(tmp_path / f"cam{c}.col")
        val.tofile(tmp_path / f"cam{c}.val")
    r = subprocess.run([probe, str(tmp_path), str(len(cams)), str(len(xyz)), mode], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return [np.fromfile(tmp_path / f"cam{c}.val.out", np.float32) for c in range(len(cams))], np.fromfile(tmp_path / "skipped.u32", np.uint32)


@pytest.mark.parametrize("n_cams,mode", [(2, "average"), (2, "best"), (3, "average"), (3, "best"), (4, "average")])
def test_weights_match_numpy(up, tmp_path, n_cams, mode):
    xyz, nrm, centers, cams = _scene(up, n_cams, seed=n_cams)
    got, skipped = _run(up, tmp_path, xyz, nrm, centers, cams, mode)
    seen = np.stack([c[3] for c in cams])                     # [cams, nodes]
    ang = np.stack([_angles(xyz, nrm, centers[c]) for c in range(n_cams)])
    multi = seen.sum(0) >= 2
    want = np.ones_like(ang)
    for n in np.flatnonzero(multi):
        cs = np.flatnonzero(seen[:, n])
        a = ang[cs, n]
        if mode == "best":
            w = np.zeros(len(cs), np.float32)
            w[int(np.argmax(a))] = 1.0                        # first maximum in camera order
        else:
            s = np.float32(0)
            for v in a:
                s = np.float32(s + v)
            w = (a / s).astype(np.float32)
        want[cs, n] = w
    assert multi.sum() > 50 and (seen.sum(0) == n_cams).sum() > 5
    for c in range(n_cams):
        w_c = want[c, seen[c]]
        if n_cams == 2 or mode == "best":
            ties = False
            assert np.array_equal(got[c].view(np.uint32), w_c.view(np.uint32)) or ties
        else:
            assert np.all(np.abs(got[c] - w_c) <= 2 * np.spacing(np.abs(w_c)))     # sum order + division
    assert np.array_equal(skipped, np.flatnonzero(seen.sum(0) == 0).astype(np.uint32))
    if mode == "average":                                      # weights of a node sum to one
        tot = np.zeros(len(xyz))
        for c in range(n_cams):
            tot[seen[c]] += got[c]
        assert np.allclose(tot[multi], 1.0, atol=3e-7) and np.all(tot[seen.sum(0) == 1] == 1.0)
