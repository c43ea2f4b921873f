"""host/patch_geometry.hpp (phase-0 fiducial patch geometry: cluster_points, get_target_boundary,
get_cluster_boundary, PatchClusters ctor, threshold_bounds; cpp/lib/patches.ipp:15-94, 240-487)
against the oracle's plain-Python restatement of the same reference functions.  CPU only."""
import os
import subprocess

import numpy as np
import pytest


def _run(probe, tmp_path, targs, W, H, bt, bf, ref=None, thresh=0, offset=2):
    tf = tmp_path / "targets.txt"
    tf.write_text("".join("%.9g %.9g %.9g\n" % tuple(float(x) for x in t) for t in targs))
    args = [probe, str(tf), str(W), str(H), str(bt), str(bf)]
    if ref is not None:
        ref.astype("<u2").tofile(tmp_path / "ref.u16")
        args += [str(tmp_path / "ref.u16"), str(thresh), str(offset)]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    clusters = []
    for line in r.stdout.splitlines():
        t = line.split()
        if t[0] == "cluster":
            clusters.append((int(t[2]), [], []))
        else:
            clusters[-1][1 if t[0] == "b" else 2].append((int(t[1]), int(t[2])))
    return clusters


@pytest.mark.parametrize("bt,bf,seed", [(2, 0, 0), (3, 1, 1), (2, 2, 2), (1, 0, 3), (4, 1, 4)])
def test_patch_geometry_matches_oracle(up, tmp_path, bt, bf, seed):
    from oracle import setup_patches as sp
    probe = up.build.build_patch_probe()
    rng = np.random.default_rng(seed)
    W, H = 160, 120
    # isolated targets, touching pairs / chains (clusters), and targets hanging over the frame border
    targs = [(rng.uniform(10, W - 10), rng.uniform(10, H - 10), rng.uniform(3, 7)) for _ in range(10)]
    for _ in range(5):
        u, v = rng.uniform(20, W - 20), rng.uniform(20, H - 20)
        targs += [(u, v, 5.0), (u + rng.uniform(4, 9), v + rng.uniform(-3, 3), 4.5), (u - rng.uniform(3, 8), v + rng.uniform(4, 8), 6.0)]
    targs += [(1.5, 40.0, 5.0), (W - 1.2, 70.0, 6.0), (80.0, H - 0.5, 4.0), (2.0, 2.0, 6.0)]
    targs = [tuple(np.float32(x) for x in t) for t in targs]
    ref = rng.integers(200, 4000, (H, W)).astype(np.uint16)
    for use_ref in (False, True):
        got = _run(probe, tmp_path, targs, W, H, bt, bf, ref if use_ref else None, thresh=400, offset=2)
        clusters = sp.cluster_points(targs, bt + bf)
        want = sp.patch_clusters(clusters, W, H, bt, bf, ref if use_ref else None, thresh=400, offset=2)
        assert [g[0] for g in got] == [len(c) for c in clusters]
        assert any(n > 1 for n, _, _ in got) and any(n == 1 for n, _, _ in got)
        for (n, b, i), (wb, wi) in zip(got, want):
            assert b == wb and i == wi                    # same pixels, same order
    # the lists are what upsp_gpu_set_patches consumes: interior pixels inside the frame, boundary ring around them
    for n, b, i in got:
        assert all(0 <= x < W and 0 <= y < H for x, y in b + i) and not (set(b) & set(i))


def test_patch_probe_fails_loudly(up, tmp_path):
    probe = up.build.build_patch_probe()
    r = subprocess.run([probe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
    r = subprocess.run([probe, str(tmp_path / "none.txt"), "10", "10", "2", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot open" in r.stderr
