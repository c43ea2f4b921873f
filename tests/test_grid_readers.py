"""host/grid_readers.hpp: the unstructured .tri grid reader (TriModel_::load_grid,
cpp/lib/TriModel.ipp:115-225) and the node normals (calcNormals, :1428-1506).  Known answers come
from the reference's own unit tests (cpp/test/test_trimodel.cpp:26-100: node / face / component
counts of its sphere fixtures); the arrays are checked against a numpy reading of the same
Fortran-unformatted records.  CPU only."""
import os
import struct
import subprocess

import numpy as np
import pytest


def write_tri(path, xyz, tri, comps=None):
    """Cart3D-style unformatted .tri: {n_node, n_tri}, xyz float32, 1-based faces int32[, components]"""
    rec = lambda payload: struct.pack("<i", len(payload)) + payload + struct.pack("<i", len(payload))
    with open(path, "wb") as f:
        f.write(rec(struct.pack("<ii", len(xyz), len(tri))))
        f.write(rec(np.ascontiguousarray(xyz, "<f4").tobytes()))
        f.write(rec(np.ascontiguousarray(tri + 1, "<i4").tobytes()))
        if comps is not None:
            f.write(rec(np.ascontiguousarray(comps, "<i4").tobytes()))


def run_probe(probe, path, prefix=None):
    r = subprocess.run([probe, str(path)] + ([str(prefix)] if prefix else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return {l.split()[0]: int(l.split()[1]) for l in r.stdout.splitlines()}


def numpy_normals(xyz, tri):
    """calcNormals in float32, faces accumulated in ascending order"""
    xyz = xyz.astype(np.float32)
    p0, p1, p2 = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    u, v = p2 - p1, p0 - p1
    n = np.stack([u[:, 1] * v[:, 2] - v[:, 1] * u[:, 2], v[:, 0] * u[:, 2] - u[:, 0] * v[:, 2],
                  u[:, 0] * v[:, 1] - v[:, 0] * u[:, 1]], 1).astype(np.float32)
    m = np.sqrt((n.astype(np.float64) ** 2).sum(1)).astype(np.float32)
    n = np.where(m[:, None] != 0, n / np.where(m == 0, 1, m)[:, None], n).astype(np.float32)
    acc = np.zeros_like(xyz)
    for t in range(len(tri)):                     # ordered accumulation, float32
        for k in range(3):
            acc[tri[t, k]] = acc[tri[t, k]] + n[t]
    m = np.sqrt((acc.astype(np.float64) ** 2).sum(1)).astype(np.float32)
    return np.where(m[:, None] != 0, acc / np.where(m == 0, 1, m)[:, None], acc).astype(np.float32)


@pytest.mark.parametrize("with_comps", [False, True])
def test_tri_reader_and_normals(up, tmp_path, with_comps):
    probe = up.build.build_grid_probe()
    xyz, _, tri = up.synth.make_sphere_mesh(10, 20, 3.0, (0.5, -1.0, 2.0), bump=0.1, seed=4)
    xyz = np.concatenate([xyz, [[9.0, 9.0, 9.0]]]).astype(np.float32)        # a node no triangle uses
    comps = (np.arange(len(tri)) % 3 + 1).astype(np.int32) if with_comps else None
    write_tri(tmp_path / "g.tri", xyz, tri, comps)
    info = run_probe(probe, tmp_path / "g.tri", tmp_path / "d")
    assert info == dict(n_nodes=len(xyz), n_tris=len(tri), n_comps=3 if with_comps else 1, has_comps=int(with_comps))
    assert np.array_equal(np.fromfile(tmp_path / "d.xyz", np.float32).reshape(-1, 3), xyz)
    assert np.array_equal(np.fromfile(tmp_path / "d.tri", np.int32).reshape(-1, 3), tri)
    if with_comps:
        assert np.array_equal(np.fromfile(tmp_path / "d.comp", np.int32), comps)
    nrm = np.fromfile(tmp_path / "d.nrm", np.float32).reshape(-1, 3)
    want = numpy_normals(xyz, tri)
    assert np.array_equal(nrm.view(np.uint32), want.view(np.uint32))
    assert np.all(nrm[-1] == 0) and np.allclose(np.linalg.norm(nrm[:-1], axis=1), 1.0, atol=1e-6)
    centre = np.float32([0.5, -1.0, 2.0])
    assert np.all(np.einsum("ij,ij->i", nrm[:-1], xyz[:-1] - centre) > 0)    # outward for this winding


def test_reference_sphere_fixtures(up):
    """cpp/test/test_trimodel.cpp:64-70, 86-92, 26-32 (build container only)"""
    base = "/root/reference/cpp/test/inputs/"
    if not os.path.exists(base + "sphere_unf_single.tri"):
        pytest.skip("reference fixtures not present on this machine")
    probe = up.build.build_grid_probe()
    assert run_probe(probe, base + "sphere_unf_single.tri") == dict(n_nodes=594, n_tris=1024, n_comps=1, has_comps=1)
    assert run_probe(probe, base + "sphere_unf_single.i.tri") == dict(n_nodes=514, n_tris=1024, n_comps=1, has_comps=1)
    multi = run_probe(probe, base + "sphere_unf_multi.i.tri")
    assert (multi["n_nodes"], multi["n_tris"], multi["n_comps"]) == (514, 1024, 6)


def test_tri_reader_errors(up, tmp_path):
    probe = up.build.build_grid_probe()
    r = subprocess.run([probe, str(tmp_path / "none.tri")], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot open tri grid file" in r.stderr
    xyz, _, tri = up.synth.make_sphere_mesh(4, 8)
    write_tri(tmp_path / "g.tri", xyz, tri)
    raw = (tmp_path / "g.tri").read_bytes()
    (tmp_path / "cut.tri").write_bytes(raw[:len(raw) // 2])
    r = subprocess.run([probe, str(tmp_path / "cut.tri")], capture_output=True, text=True)
    assert r.returncode == 1 and "inconsistent number of" in r.stderr
    (tmp_path / "bad.tri").write_bytes(struct.pack("<i", 12) + raw[4:])
    r = subprocess.run([probe, str(tmp_path / "bad.tri")], capture_output=True, text=True)
    assert r.returncode == 1 and "Unable to read tri grid file" in r.stderr


# ---------------------------------------------------------------- structured plot3d grids
def write_p3d(path, zones, dp=False, big=False, iblank=False, multi=None):
    """zones: list of (x, y, z) arrays of shape (k, j, i).  Unformatted plot3d as plot3d.h:30-60."""
    e = ">" if big else "<"
    ft = e + ("f8" if dp else "f4")
    rec = lambda payload: struct.pack(e + "i", len(payload)) + payload + struct.pack(e + "i", len(payload))
    multi = len(zones) > 1 if multi is None else multi
    with open(path, "wb") as f:
        if multi:
            f.write(rec(struct.pack(e + "i", len(zones))))
        f.write(rec(b"".join(struct.pack(e + "iii", *z[0].shape[::-1]) for z in zones)))
        for x, y, z in zones:
            payload = b"".join(np.ascontiguousarray(a, ft).tobytes() for a in (x, y, z))
            if iblank:
                payload += np.ones(x.size, e + "i4").tobytes()
            f.write(rec(payload))


def _zones(seed, shapes):
    rng = np.random.default_rng(seed)
    return [tuple(rng.normal(0, 5, s).astype(np.float32) for _ in range(3)) for s in shapes]


@pytest.mark.parametrize("shapes", [[(1, 7, 9)], [(1, 5, 4), (2, 3, 6), (1, 1, 8)]], ids=["single", "multi"])
def test_plot3d_reader_all_variants_round_trip(up, tmp_path, shapes):
    """The reference's own test (cpp/test/test_plot3d.cpp ReadWriteGridFiles): whatever the variant
    read (precision, endianness, IBLANK), writing it back gives the little-endian file of the
    requested precision, byte for byte."""
    probe = up.build.build_grid_probe()
    zones = _zones(len(shapes), shapes)
    write_p3d(tmp_path / "le_sp.x", zones)
    write_p3d(tmp_path / "le_dp.x", zones, dp=True)
    for dp in (False, True):
        for big in (False, True):
            for ib in (False, True):
                src = tmp_path / f"v_{dp}_{big}_{ib}.x"
                write_p3d(src, zones, dp=dp, big=big, iblank=ib)
                for out_dp in (False, True):
                    out = tmp_path / "o.x"
                    r = subprocess.run([probe, str(src), "dp" if out_dp else "sp", str(out)], capture_output=True, text=True)
                    assert r.returncode == 0, r.stderr
                    assert out.read_bytes() == (tmp_path / ("le_dp.x" if out_dp else "le_sp.x")).read_bytes()
                    lines = r.stdout.splitlines()
                    assert lines[0] == f"n_zones {len(shapes)}" and lines[1] == f"n_points {sum(int(np.prod(s)) for s in shapes)}"
                    assert [tuple(map(int, l.split()[2:])) for l in lines[2:]] == [s[::-1] for s in shapes]


def test_plot3d_reference_fixtures(up, tmp_path):
    """cpp/test/test_plot3d.cpp:36-118 on the reference's fixtures, plus the reference's Python reader
    (python/upsp/processing/plot3d.py read_p3d_grid) on the single-precision multi-zone file."""
    base = "/root/reference/cpp/test/inputs/sphere_unf_"
    if not os.path.exists(base + "multi_integration_sp.x"):
        pytest.skip("reference fixtures not present on this machine")
    probe = up.build.build_grid_probe()
    for kind in ("single", "multi"):
        variants = ["sp", "dp", "sp_bigend", "dp_bigend"] + (["sp_iblank", "dp_iblank"] if kind == "multi" else [])
        for v in variants:
            for prec in ("sp", "dp"):
                out = tmp_path / "o.x"
                r = subprocess.run([probe, f"{base}{kind}_integration_{v}.x", prec, str(out)], capture_output=True, text=True)
                assert r.returncode == 0, r.stderr
                assert out.read_bytes() == open(f"{base}{kind}_integration_{prec}.x", "rb").read()
    import sys
    sys.path.insert(0, "/root/reference/python")
    try:
        from upsp.processing import plot3d
    except Exception:
        return
    if not hasattr(np, "product"):
        np.product = np.prod          # the reference predates numpy 2
    g = plot3d.read_p3d_grid(base + "multi_integration_sp.x")
    out = tmp_path / "m.x"
    subprocess.run([probe, base + "multi_integration_sp.x", "sp", str(out)], check=True, capture_output=True)
    raw = out.read_bytes()
    nz = struct.unpack("<i", raw[4:8])[0]
    assert nz == len(g.sz) == 4
    pos, xs = 12 + 4 + 12 * nz + 4, []
    for z in range(nz):
        n = int(np.prod(g.sz[z]))
        pos += 4
        xs.append(np.frombuffer(raw, "<f4", 3 * n, pos)[:n])
        pos += 12 * n + 4
    assert np.array_equal(np.concatenate(xs), np.asarray(g.x, np.float32))


def test_plot3d_reader_errors(up, tmp_path):
    probe = up.build.build_grid_probe()
    r = subprocess.run([probe, str(tmp_path / "none.x"), "sp"], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot open plot3d grid file" in r.stderr
    (tmp_path / "junk.x").write_bytes(struct.pack("<i", 77) + bytes(100))
    r = subprocess.run([probe, str(tmp_path / "junk.x"), "sp"], capture_output=True, text=True)
    assert r.returncode == 1 and "bad header record" in r.stderr
    write_p3d(tmp_path / "ok.x", _zones(1, [(1, 3, 4)]))
    raw = (tmp_path / "ok.x").read_bytes()
    (tmp_path / "cut.x").write_bytes(raw[:-20])
    r = subprocess.run([probe, str(tmp_path / "cut.x"), "sp"], capture_output=True, text=True)
    assert r.returncode == 1 and "truncated zone data" in r.stderr


def numpy_intersect(xyz, tri, comps):
    """TriModel_::intersect_grid (cpp/lib/TriModel.ipp:941-1118) restated: lowest index of every group of identical nodes,
    collapsed triangles dropped, orphans between consecutive removed nodes removed, nodes renumbered in order"""
    n = len(xyz)
    groups = {}
    for i, p in enumerate(xyz):
        groups.setdefault(tuple(float(v) + 0.0 for v in p), []).append(i)           # + 0.0: -0.0 == 0.0
    rep = np.arange(n)
    for g in groups.values():
        rep[g] = g[0]
    removed = np.flatnonzero(rep != np.arange(n))
    if len(removed) == 0:
        return xyz, tri, comps, 0
    t = rep[tri]
    keep = (t[:, 0] != t[:, 1]) & (t[:, 0] != t[:, 2]) & (t[:, 1] != t[:, 2])
    t, comps = t[keep], comps[keep]
    used = np.zeros(n, bool)
    used[t.ravel()] = True
    gone = np.zeros(n, bool)
    gone[removed] = True
    orphans = 0
    for i, r in enumerate(removed):
        end = removed[i + 1] if i + 1 < len(removed) else len(removed)
        for k in range(r + 1, end):
            if not used[k] and not gone[k]:
                gone[k] = True
                orphans += 1
    new = np.cumsum(~gone) - 1
    return xyz[~gone], new[t].astype(np.int32), comps, sum(len(g) > 1 for g in groups.values()) + orphans


def test_intersect_grid_reproduces_reference_fixtures(up, tmp_path):
    """psp_process loads a .tri grid with intersect = true: the reference's un-intersected sphere fixtures must become its own
    intersected ones (*.i.tri) node for node and triangle for triangle (known answers of cpp/test/test_trimodel.cpp:62-79:
    594 -> 514 nodes, 1024 faces, 1 / 6 components; intersecting an intersected grid changes nothing)."""
    base = "/root/reference/cpp/test/inputs/sphere_unf_"
    if not os.path.exists(base + "single.tri"):
        pytest.skip("reference fixtures not present on this machine")
    probe = up.build.build_grid_probe()
    for kind, ncomp in (("single", 1), ("multi", 6)):
        r = subprocess.run([probe, base + kind + ".tri", str(tmp_path / "a"), "intersect"], capture_output=True, text=True)
        info = {l.split()[0]: int(l.split()[1]) for l in r.stdout.splitlines()}
        assert info["n_nodes"] == 514 and info["n_tris"] == 1024 and info["n_comps"] == ncomp and info["non_unique"] == 80
        r = subprocess.run([probe, base + kind + ".i.tri", str(tmp_path / "b"), "intersect"], capture_output=True, text=True)
        info = {l.split()[0]: int(l.split()[1]) for l in r.stdout.splitlines()}
        assert info["n_nodes"] == 514 and info["n_tris"] == 1024 and info["non_unique"] == 0
        for ext, dt in ((".xyz", np.float32), (".tri", np.int32), (".comp", np.int32)):
            assert np.array_equal(np.fromfile(str(tmp_path / "a") + ext, dt), np.fromfile(str(tmp_path / "b") + ext, dt)), (kind, ext)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_intersect_grid_matches_restatement(up, tmp_path, seed):
    probe = up.build.build_grid_probe()
    rng = np.random.default_rng(seed)
    xyz, _, tri = up.synth.make_sphere_mesh(8, 12, 2.0, (0, 0, 0), seed=seed)
    xyz = xyz.astype(np.float32)
    n0 = len(xyz)
    dup_of = rng.choice(n0, 25, replace=False)                              # 25 duplicated nodes, some twice, spread over the list
    extra = np.concatenate([xyz[dup_of], xyz[dup_of[:6]]])
    pos = np.sort(rng.choice(n0, len(extra), replace=False))
    xyz2 = np.insert(xyz, pos, extra, axis=0)
    shift = np.arange(n0) + np.searchsorted(pos, np.arange(n0), side="right")   # new index of every original node
    tri2 = shift[tri].astype(np.int32)
    new_ids = pos + np.arange(len(pos))
    for k in range(0, len(new_ids), 2):                                     # half of the copies take over a triangle corner
        orig = shift[np.concatenate([dup_of, dup_of[:6]])[k]]
        hits = np.argwhere(tri2 == orig)
        if len(hits):
            tri2[tuple(hits[0])] = new_ids[k]
    a, b = shift[dup_of[0]], new_ids[0]
    tri2 = np.concatenate([tri2, [[a, b, shift[tri[0, 0]]]]]).astype(np.int32)  # a triangle that collapses (two identical corners)
    xyz2 = np.concatenate([xyz2, [[9, 9, 9]]]).astype(np.float32)           # an orphan after the last removed node: stays
    xyz2[3] = np.where(xyz2[3] == 0, -0.0, xyz2[3])
    comps = (np.arange(len(tri2)) % 4 + 1).astype(np.int32)
    write_tri(tmp_path / "g.tri", xyz2, tri2, comps)
    r = subprocess.run([probe, str(tmp_path / "g.tri"), str(tmp_path / "d"), "intersect"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = {l.split()[0]: int(l.split()[1]) for l in r.stdout.splitlines()}
    wx, wt, wc, overlap = numpy_intersect(xyz2, tri2, comps)
    assert info["non_unique"] == len(xyz2) - len(wx) >= 25 and info["unique_overlapping"] == overlap
    assert np.array_equal(np.fromfile(tmp_path / "d.xyz", np.float32).reshape(-1, 3), wx)
    assert np.array_equal(np.fromfile(tmp_path / "d.tri", np.int32).reshape(-1, 3), wt) and len(wt) < len(tri2)
    assert np.array_equal(np.fromfile(tmp_path / "d.comp", np.int32), wc)


def numpy_area_weighted_normals(xyz, tri):
    """TriModel_::Node::get_normal (TriModel.ipp:1570-1590) with upsp::normal / upsp::area (models.ipp:137-182), float32
    arithmetic in the reference's order, faces accumulated in ascending index"""
    f32, f64 = np.float32, np.float64
    xyz = xyz.astype(f32)
    acc = np.zeros_like(xyz)
    nd = lambda v: np.sqrt((v.astype(f64) ** 2).sum())
    for t in tri:
        p0, p1, p2 = xyz[t[0]], xyz[t[1]], xyz[t[2]]
        a, b = (p2 - p1).astype(f32), (p0 - p1).astype(f32)
        n = np.array([f32(f32(a[1] * b[2]) - f32(a[2] * b[1])), f32(f32(a[2] * b[0]) - f32(a[0] * b[2])), f32(f32(a[0] * b[1]) - f32(a[1] * b[0]))], f32)
        nn = nd(n)
        if f32(nn) != 0:
            n = (n.astype(f64) / nn).astype(f32)
        ea, eb, ec = f32(nd((p1 - p0).astype(f32))), f32(nd((p2 - p1).astype(f32))), f32(nd((p2 - p0).astype(f32)))
        if eb > ea:
            ea, eb = eb, ea
        if ec > ea:
            ea, eb, ec = ec, ea, eb
        elif ec > eb:
            eb, ec = ec, eb
        pos_neg = f32(abs(f32(ec - f32(ea - eb))))
        prod = f32(f32(f32(f32(ea + f32(eb + ec)) * pos_neg) * f32(ec + f32(ea - eb))) * f32(ea + f32(eb - ec)))
        area = f32(0.25 * float(np.sqrt(prod, dtype=f32)))
        for k in t:
            acc[k] = (acc[k] + (n * area).astype(f32)).astype(f32)
    out = acc.copy()
    for i, v in enumerate(acc):
        m = nd(v)
        if m != 0:
            out[i] = (v.astype(f64) / m).astype(f32)
    return out


def test_area_weighted_node_normals(up, tmp_path):
    """what the reference's camera weights and target diameters take as the node normal of an unstructured model"""
    probe = up.build.build_grid_probe()
    xyz, _, tri = up.synth.make_sphere_mesh(9, 14, 3.0, (0.5, -1.0, 2.0), bump=0.15, seed=8)
    xyz = np.concatenate([xyz, [[7.0, 7.0, 7.0]]]).astype(np.float32)        # a node without triangles: zero normal
    write_tri(tmp_path / "g.tri", xyz, tri, np.ones(len(tri), np.int32))
    run_probe(probe, tmp_path / "g.tri", tmp_path / "d")
    got = np.fromfile(tmp_path / "d.nrmw", np.float32).reshape(-1, 3)
    want = numpy_area_weighted_normals(xyz, tri)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert not np.any(got[-1]) and np.allclose(np.linalg.norm(got[:-1], axis=1), 1, atol=1e-6)
    plain = np.fromfile(tmp_path / "d.nrm", np.float32).reshape(-1, 3)
    assert not np.array_equal(got, plain) and np.abs(got - plain).max() < 0.3      # close to, but not, the unweighted normals
