"""host/interpolation.hpp: upsp::interpolate (cpp/lib/interpolation.ipp:17-67) as phase 2 uses it for wind-on runs on
an unstructured model (psp_process.cpp:2338-2345, 2371-2378): k = 10 nearest valid nodes of the steady grid, inverse
distance squared.  Held against an exhaustive numpy search with the same float arithmetic.  CPU only."""
import subprocess

import numpy as np
import pytest

from oracle import p3d_overlap
from test_p3d_model import seam_grid, write_p3d


def numpy_interpolate(xyz_in, valid, data, pts, k, p=2.0):
    f32 = np.float32
    out = np.zeros(len(pts), f32)
    idx_valid = np.nonzero(valid)[0]
    xin = xyz_in.astype(f32)
    for i, pt in enumerate(pts.astype(f32)):
        d2 = ((xin[idx_valid].astype(np.float64) - pt.astype(np.float64)) ** 2).sum(1)
        order = np.lexsort((idx_valid, d2))[:k]                      # nearest first, ties by index
        total = f32(0)
        for n in idx_valid[order]:
            d = xin[n] - pt                                          # float32 differences
            dist = f32(np.sqrt((d.astype(np.float64) ** 2).sum()))
            if dist == 0:
                total, out[i] = f32(1), data[n]
                break
            w = f32(1.0 / float(dist) ** p)
            out[i] = f32(out[i] + f32(data[n] * w))
            total = f32(total + w)
        out[i] = f32(out[i] / total)
    return out


@pytest.mark.parametrize("k", [1, 10])
def test_interpolate_matches_exhaustive_search(up, tmp_path, k):
    probe = up.build.build_inputs_probe()
    zones = seam_grid(3)
    write_p3d(tmp_path / "g.x", zones)
    xyz = np.concatenate([z[2] for z in zones]).astype(np.float32)
    sizes = [(z[0], z[1]) for z in zones]
    overlap, _, _ = p3d_overlap.identify_overlap(xyz, sizes, 1e-3)
    valid = np.array([p3d_overlap.get_low_nidx(overlap, n) == n for n in range(len(xyz))])
    rng = np.random.default_rng(5)
    data = rng.normal(0, 1, len(xyz)).astype(np.float32)
    lo, hi = xyz.min(0), xyz.max(0)
    pts = np.concatenate([rng.uniform(lo - 0.3, hi + 0.3, (300, 3)),            # around and inside the grid
                          xyz[[0, 17, 60]],                                     # exact hits (node 60 may be a seam node)
                          rng.uniform(lo - 40, hi + 40, (20, 3))]).astype(np.float32)   # far outside the box
    data.tofile(tmp_path / "d.f32")
    pts.tofile(tmp_path / "p.f32")
    r = subprocess.run([probe, "interp", str(tmp_path / "g.x"), "0.001", str(tmp_path / "d.f32"), str(tmp_path / "p.f32"), str(k),
                        str(tmp_path / "o.f32")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(tmp_path / "o.f32", np.float32)
    want = numpy_interpolate(xyz, valid, data, pts, k)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert got[300] == data[0] and got[301] == data[17]
