"""host/p3d_model.hpp: the structured surface model around row a8 (P3DModel_::identifyOverlap,
cpp/lib/P3DModel.ipp:893-1127; get_low_nidx :1346-1354; adjust_solution :144-157 -> the static src_index of
upsp_gpu_set_overlap_remap; extract_tris :234-317; calcNormals :1357-1654).  Known answers are the reference's own
unit test (cpp/test/test_p3dmodel.cpp:27-64, 166-240 on the fixture of cpp/test/test_grid_utils.cpp:49-123);
random multi-zone grids with seams and a wrapped zone are held against the plain-Python restatement
oracle/p3d_overlap.py.  CPU only."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import p3d_overlap


def write_p3d(path, zones):
    """multi-zone unformatted plot3d, single precision; zones = [(J, K, xyz[K*J, 3])]"""
    rec = lambda b: struct.pack("<i", len(b)) + b + struct.pack("<i", len(b))
    with open(path, "wb") as f:
        f.write(rec(struct.pack("<i", len(zones))))
        f.write(rec(b"".join(struct.pack("<3i", J, K, 1) for J, K, _ in zones)))
        for J, K, xyz in zones:
            f.write(rec(np.ascontiguousarray(xyz.T, "<f4").tobytes()))


def run_probe(probe, grid, tol, prefix):
    r = subprocess.run([probe, "overlap", str(grid), repr(float(tol)), str(prefix)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = {}
    for line in r.stdout.splitlines():
        f = line.split()
        info.setdefault(f[0], []).append([int(v) for v in f[1:]])
    return (info, np.fromfile(str(prefix) + ".src", np.int32), np.fromfile(str(prefix) + ".pairs", np.int32).reshape(-1, 2),
            np.fromfile(str(prefix) + ".trinodes", np.int32), np.fromfile(str(prefix) + ".nrm", np.float32).reshape(-1, 3))


def reference_fixture(offset=0.0):
    """create_single_structgrid + add_zones_structgrid, cpp/test/test_grid_utils.cpp:49-123"""
    f32 = np.float32
    z0 = np.array([[j, k, 0] for k in range(5) for j in range(4)], f32)
    z1 = np.array([[f32(j + 3) + f32(offset), k, 0] for k in range(4) for j in range(3)], f32)
    z2 = np.array([[f32(6.0 - k + f32(offset)), f32(j - 4.0 - f32(offset)), 0] for k in range(4) for j in range(5)], f32)
    return [(4, 5, z0), (3, 4, z1), (5, 4, z2)]


@pytest.fixture(scope="module")
def probe(up):
    return up.build.build_inputs_probe()


def test_reference_known_answers(probe, tmp_path):
    write_p3d(tmp_path / "single.x", reference_fixture()[:1])
    info, src, pairs, tri, nrm = run_probe(probe, tmp_path / "single.x", 1e-10, tmp_path / "s")
    assert info["n_nodes"] == [[20]] and info["n_faces"] == [[12]] and info["n_zones"] == [[1]] and info["zone"] == [[0, 4, 5, 0]]
    assert len(pairs) == 0 and np.array_equal(src, np.arange(20)) and np.array_equal(nrm, np.tile(np.float32([0, 0, 1]), (20, 1)))

    write_p3d(tmp_path / "multi.x", reference_fixture())
    info, src, pairs, tri, nrm = run_probe(probe, tmp_path / "multi.x", 1e-10, tmp_path / "m")
    assert info["n_nodes"] == [[52]] and info["n_faces"] == [[12 + 6 + 12]] and info["n_zones"] == [[3]]
    assert info["zone"] == [[0, 4, 5, 0], [1, 3, 4, 20], [2, 5, 4, 32]]
    superceded = np.nonzero(src != np.arange(52))[0]
    assert sorted(superceded) == [20, 23, 26, 29, 41, 46, 51]              # ExactNodeOverlap
    assert info["n_superceded"] == [[7]] and info["n_vert"] == [[52 - 7]]   # NodeIterator: 52 - 7 nodes are visited
    assert np.array_equal(nrm, np.tile(np.float32([0, 0, 1]), (52, 1)))


@pytest.mark.parametrize("tol, merged", [(0.09, 0), (0.100001, 7)])
def test_reference_tolerance_cases(probe, tmp_path, tol, merged):
    write_p3d(tmp_path / "g.x", reference_fixture(offset=0.1))              # TolNodeOverlap
    info, src, *_ = run_probe(probe, tmp_path / "g.x", tol, tmp_path / "g")
    assert int((src != np.arange(52)).sum()) == merged and info["n_vert"] == [[52 - merged]]


def seam_grid(seed):
    """three zones: a patch, a neighbour sharing one edge (perturbed within / beyond the tolerance), and a closed
    cylinder whose first and last j-columns coincide (a wrapped zone) and whose rim touches the patch"""
    rng = np.random.default_rng(seed)
    J0, K0 = 7, 6
    a = np.array([[j, k, 0.1 * np.sin(j + k)] for k in range(K0) for j in range(J0)], np.float32)
    J1, K1 = 5, 6
    b = np.array([[J0 - 1 + j, k, 0.1 * np.sin(J0 - 1 + j + k)] for k in range(K1) for j in range(J1)], np.float32)
    b[::J1] = a[J0 - 1::J0]                                       # shared edge, bit-identical ...
    b[0] += np.float32([4e-4, 0, 0])                              # ... one node inside the tolerance
    b[J1] += np.float32([0, 3e-3, 0])                             # ... one node beyond it
    J2, K2 = 9, 4
    th = np.linspace(0, 2 * np.pi, J2)
    c = np.array([[3 + np.cos(t), -2 - k, np.sin(t)] for k in range(K2) for t in th], np.float32)
    c[J2 - 1::J2] = c[::J2]                                       # closed: j = 0 and j = J-1 coincide
    c[3] = a[2]                                                   # rim node on the patch's lower edge
    extra = rng.normal(0, 1e-5, c.shape).astype(np.float32)
    extra[::J2] = extra[J2 - 1::J2] = 0
    extra[3] = 0
    return [(J0, K0, a), (J1, K1, b), (J2, K2, c + extra)]


@pytest.mark.parametrize("seed", [0, 1])
def test_against_restatement(probe, tmp_path, seed):
    zones = seam_grid(seed)
    write_p3d(tmp_path / "g.x", zones)
    tol = 1e-3
    info, src, pairs, tri, nrm = run_probe(probe, tmp_path / "g.x", tol, tmp_path / "g")
    xyz = np.concatenate([z[2] for z in zones])
    sizes = [(z[0], z[1]) for z in zones]
    want, nonuniq, uniq = p3d_overlap.identify_overlap(xyz, sizes, tol)
    want_pairs = np.array([(a, b) for a in sorted(want) for b in want[a]], np.int32).reshape(-1, 2)
    assert len(want_pairs) == 20 and np.array_equal(pairs, want_pairs)
    assert info["n_vert"] == [[len(xyz) - nonuniq + uniq]]
    assert np.array_equal(src, p3d_overlap.adjust_solution(want, np.arange(len(xyz))))
    assert np.array_equal(tri, p3d_overlap.extract_tri_nodes(sizes))
    # the wrapped zone merges its seam column, the perturbed nodes behave as their distance says
    s2 = sum(j * k for j, k in sizes[:2])
    assert src[s2 + 8] == s2 and src[42] == 6 and src[42 + 5] == 42 + 5
    # seam nodes share one normal: both sides sum the faces of both zones
    for a, bs in want.items():
        for b in bs:
            if len(bs) == 1 and len(want[b]) == 1:
                assert np.allclose(nrm[a], nrm[b], atol=2e-6)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1, atol=1e-5)


def test_reference_volume_grid(probe, tmp_path):
    path = "/root/reference/test/data/fml_tc3_volume.grid"
    if not os.path.exists(path):
        pytest.skip("reference fixture not present on this machine")
    info, src, pairs, tri, nrm = run_probe(probe, path, 1e-3, tmp_path / "v")
    assert info["n_nodes"] == [[309062]] and info["n_zones"] == [[14]] and len(tri) == 6 * info["n_faces"][0][0]
    raw = open(path, "rb").read()                             # multi-zone, single precision, little endian
    nz = struct.unpack_from("<i", raw, 4)[0]
    dims = np.frombuffer(raw, "<i4", 3 * nz, 16).reshape(nz, 3)
    off, xyz = 16 + 12 * nz + 4, []
    for j, k, l in dims:
        n = int(j * k * l)
        xyz.append(np.frombuffer(raw, "<f4", 3 * n, off + 4).reshape(3, n).T)
        off += 12 * n + 8
    assert off == len(raw)
    xyz = np.concatenate(xyz)
    sizes = [(int(d[0]), int(d[1])) for d in dims]
    want, nonuniq, uniq = p3d_overlap.identify_overlap(xyz, sizes, 1e-3)
    want_pairs = np.array([(a, b) for a in sorted(want) for b in want[a]], np.int32).reshape(-1, 2)
    assert np.array_equal(pairs, want_pairs)
    assert info["n_vert"] == [[len(xyz) - nonuniq + uniq]]
    assert np.array_equal(src, p3d_overlap.adjust_solution(want, np.arange(len(xyz))))
