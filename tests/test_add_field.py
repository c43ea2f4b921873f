"""add_field_b200: a flat [rows x extent] float file becomes a chunked dataset of an existing HDF5 file (host/h5_append.hpp,
no HDF5 library).  Reference: cpp/exec/add_field.cpp:22-128 (arguments, checks, dataset creation properties) and its batch
use docs/sphinx/quick-start.rst:125-160 (`add_field <h5> frames <pressure_transpose> <frames>`).

Checked with tests/h5min.py (the reader held against the reference's fixtures in tests/test_psp_hdf5.py) on files written by
host/psp_hdf5.hpp and on copies of the reference's own fixtures (written by the real library, group leaf K = 4, a deflated
chunked `frames` already inside), plus structural checks of what the tool appends that a lenient reader would not notice:
B-tree signatures, node types, levels, sibling links, key order, the library's end key of the chunk tree, heap offsets."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import h5min

REF_INPUTS = "/root/reference/cpp/test/inputs"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_INPUTS), reason="reference tree not present (GPU box)")
UNDEF = 0xFFFFFFFFFFFFFFFF


@pytest.fixture(scope="module")
def tool():
    import upsp_b200
    return upsp_b200.build.build_add_field_tool()


@pytest.fixture(scope="module")
def probe():
    import upsp_b200
    return upsp_b200.build.build_h5_probe()


def run(tool, *args):
    return subprocess.run([tool, *map(str, args)], capture_output=True, text=True)


def check_chunk_tree(f, path, rows, extent):
    """Every node of the dataset's chunk B-tree: type 1, levels down to 0, <= 64 entries, sibling links of each level
    chained left to right, keys ascending, a child's first key repeated by its parent, end key {rows, extent, 4} / 0 bytes."""
    b = f.b
    ds = f.root[path]
    d = ds.layout
    assert d[0] == 3 and d[1] == 2 and d[2] == 3
    addr = struct.unpack_from("<Q", d, 3)[0]
    assert struct.unpack_from("<3I", d, 11) == (1, extent, 4)
    ksz = 8 + 8 * 3
    chunks = []
    by_level = {}

    def node(a, expect_first=None):
        assert b[a:a + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
        left, right = struct.unpack_from("<QQ", b, a + 8)
        assert ntype == 1 and 1 <= used <= 64
        by_level.setdefault(level, []).append((a, left, right))
        keys, kids = [], []
        pos = a + 24
        for i in range(used + 1):
            nbytes, mask = struct.unpack_from("<II", b, pos)
            offs = struct.unpack_from("<3Q", b, pos + 8)
            assert mask == 0
            keys.append((nbytes, offs))
            if i < used:
                kids.append(struct.unpack_from("<Q", b, pos + ksz)[0])
            pos += ksz + 8
        assert [k[1] for k in keys] == sorted(k[1] for k in keys) and len(set(k[1] for k in keys)) == len(keys)
        if expect_first is not None:
            assert keys[0] == expect_first
        for i, kid in enumerate(kids):
            if level > 0:
                node(kid, keys[i])
            else:
                assert keys[i][0] == extent * 4 and keys[i][1][1:] == (0, 0)
                chunks.append((keys[i][1][0], kid))
        return keys[-1]

    end = node(addr)
    assert end == (0, (rows, extent, 4))
    assert [c[0] for c in chunks] == list(range(rows))
    for level, nodes in by_level.items():
        assert nodes[0][1] == UNDEF and nodes[-1][2] == UNDEF
        for (a0, _, r0), (a1, l1, _) in zip(nodes, nodes[1:]):
            assert r0 == a1 and l1 == a0
    assert max(by_level) + 1 == len(by_level)
    return chunks


def check_root_group(f):
    """The root group's B-tree: node type 0, symbol-table nodes at the leaves with <= 2 * leaf K entries, names ascending
    across the whole tree, every key = heap offset of the largest name to its left (0 = empty string for the first)."""
    b = f.b
    _, ohdr, _, _ = struct.unpack_from("<QQII", b, 56)
    st = [d for t, d in f._messages(ohdr) if t == 0x0011][0]
    btree, heap = struct.unpack_from("<QQ", st, 0)
    cache_type = struct.unpack_from("<I", b, 72)[0]
    if cache_type == 1:
        assert struct.unpack_from("<QQ", b, 80) == (btree, heap)
    assert b[heap:heap + 4] == b"HEAP"
    names = []

    def node(a):
        if b[a:a + 4] == b"SNOD":
            n = struct.unpack_from("<H", b, a + 6)[0]
            assert 1 <= n <= 2 * f.leaf_k
            mine = [f._heap_name(heap, struct.unpack_from("<Q", b, a + 8 + 40 * i)[0]) for i in range(n)]
            names.extend(mine)
            return mine[-1]
        assert b[a:a + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
        assert ntype == 0 and 1 <= used <= 2 * f.int_k
        pos = a + 24
        last = None
        for i in range(used):
            key, child = struct.unpack_from("<QQ", b, pos)
            if names:
                assert f._heap_name(heap, key) == names[-1]
            else:
                assert f._heap_name(heap, key) == ""
            last = node(child)
            pos += 16
        assert f._heap_name(heap, struct.unpack_from("<Q", b, pos)[0]) == last
        return last

    node(btree)
    assert names == sorted(names) and len(set(names)) == len(names)
    assert struct.unpack_from("<Q", b, 40)[0] == len(b)              # end-of-file address = file size
    return names


def snapshot(f):
    return {p: (o.shape, o.dtype, np.asarray(o.data).copy(), dict(o.attrs)) for p, o in h5min.walk(f.root) if isinstance(o, h5min.Dataset)}


def assert_same_objects(before, f):
    for p, (shape, dtype, data, attrs) in before.items():
        o = f.root[p]
        assert o.shape == shape and o.dtype == dtype and np.array_equal(np.asarray(o.data), data), p
        assert {k: list(v) for k, v in o.attrs.items()} == {k: list(v) for k, v in attrs.items()}, p


@pytest.mark.parametrize("structured", ["unstructured", "structured"])
def test_frames_added_to_a_psp_process_file(tool, probe, tmp_path, structured):
    """The documented use: psp_process's own HDF5 file + pressure_transpose -> dataset `frames` [nodes x frames]."""
    n_nodes, n_frames = 600, 40
    h5 = str(tmp_path / "out.h5")
    assert subprocess.run([probe, h5, structured, str(n_nodes)], capture_output=True).returncode == 0
    before_f = h5min.File(h5)
    before = snapshot(before_f)
    root_attrs = {k: list(v) for k, v in before_f.root.attrs.items()}
    data = np.random.default_rng(1).standard_normal((n_nodes, n_frames)).astype(np.float32)
    data[3, 5] = np.nan
    flat = str(tmp_path / "pressure_transpose")
    data.tofile(flat)
    r = run(tool, h5, "frames", flat, n_frames)
    assert r.returncode == 0, r.stdout + r.stderr
    f = h5min.File(h5)
    fr = f.root["frames"]
    assert fr.shape == (n_nodes, n_frames) and fr.dtype == np.dtype("<f4") and fr.filters == []
    assert np.array_equal(fr.data.view(np.uint32), data.view(np.uint32))
    assert_same_objects(before, f)
    assert {k: list(v) for k, v in f.root.attrs.items()} == root_attrs
    chunks = check_chunk_tree(f, "frames", n_nodes, n_frames)
    assert all(c1[1] - c0[1] == n_frames * 4 for c0, c1 in zip(chunks, chunks[1:]))     # rows lie in file order
    assert "frames" in check_root_group(f)
    # the dataset's header messages: what the library writes for the reference's creation property list
    msgs = dict((t, d) for t, d in f._messages(f.addr_of("/frames")))
    assert msgs[0x0001][:8] == bytes([1, 2, 1, 0, 0, 0, 0, 0])
    assert struct.unpack_from("<4Q", msgs[0x0001], 8) == (n_nodes, n_frames, n_nodes, n_frames)
    assert msgs[0x0005][:12] == bytes([2, 3, 2, 1, 4, 0, 0, 0, 0, 0, 0, 0])              # fill value 0.0, defined
    assert msgs[0x0004][:8] == bytes([4, 0, 0, 0, 0, 0, 0, 0])


@needs_ref
@pytest.mark.parametrize("name", ["unstruct_nodal_pencil.h5", "unstruct_nodal_pencil_trans.h5"])
def test_dataset_added_to_a_file_written_by_the_real_library(tool, tmp_path, name):
    h5 = str(tmp_path / name)
    shutil.copyfile(os.path.join(REF_INPUTS, name), h5)
    os.chmod(h5, 0o644)
    ref = h5min.File(h5)
    before = snapshot(ref)
    shape = ref.root["frames"].shape
    # the fixture's own frames (inflated by the reader), written back as a flat file under a second name
    flat = str(tmp_path / "flat")
    np.ascontiguousarray(ref.root["frames"].data).tofile(flat)
    r = run(tool, h5, "/frames_flat", flat, shape[1])
    assert r.returncode == 0, r.stdout + r.stderr
    f = h5min.File(h5)
    assert np.array_equal(f.root["frames_flat"].data, ref.root["frames"].data)
    assert_same_objects(before, f)
    assert f.root.attrs["code_version"] == ["sample scripts"]
    check_chunk_tree(f, "frames_flat", *shape)
    assert check_root_group(f) == ["Condition", "Grid", "coverage", "frames", "frames_flat"]


@needs_ref
def test_many_links_split_over_symbol_table_nodes_and_tree_levels(tool, tmp_path):
    """Group leaf K = 4, internal K = 16 in the library's files: 8 links per symbol-table node, 32 nodes per B-tree node;
    300 links need 38 symbol-table nodes and a second tree level.  Dataset of 5000 rows: three levels of chunk-tree nodes."""
    h5 = str(tmp_path / "many.h5")
    shutil.copyfile(os.path.join(REF_INPUTS, "unstruct_nodal_pencil.h5"), h5)
    os.chmod(h5, 0o644)
    before = snapshot(h5min.File(h5))
    rng = np.random.default_rng(2)
    small = rng.standard_normal((3, 5)).astype(np.float32)
    small.tofile(str(tmp_path / "small"))
    names = [f"v{(i * 7919) % 1000:03d}" for i in range(296)]
    for nm in names:
        r = run(tool, h5, nm, tmp_path / "small", 5)
        assert r.returncode == 0, r.stdout + r.stderr
    big = rng.standard_normal((5000, 6)).astype(np.float32)
    big.tofile(str(tmp_path / "big"))
    assert run(tool, h5, "big", tmp_path / "big", 6).returncode == 0
    f = h5min.File(h5)
    got = check_root_group(f)
    assert got == sorted(names + ["Condition", "Grid", "coverage", "frames", "big"])
    assert_same_objects(before, f)
    for nm in names[::37]:
        assert np.array_equal(f.root[nm].data, small)
    assert np.array_equal(f.root["big"].data, big)
    check_chunk_tree(f, "big", 5000, 6)


def test_argument_and_file_checks(tool, probe, tmp_path):
    h5 = str(tmp_path / "out.h5")
    assert subprocess.run([probe, h5, "unstructured", "50"], capture_output=True).returncode == 0
    size0 = os.path.getsize(h5)
    flat = str(tmp_path / "flat")
    np.zeros(50 * 8, np.float32).tofile(flat)
    assert run(tool, h5, "frames", flat).returncode != 0                       # three arguments
    assert run(tool, h5, "frames", flat, 0).returncode != 0                    # extent > 0
    assert run(tool, h5, "frames", flat, 7).returncode != 0                    # size not a multiple of one row
    assert run(tool, h5, "frames", str(tmp_path / "missing"), 8).returncode != 0
    assert run(tool, h5, "Grid/frames", flat, 8).returncode != 0               # root group only
    r = run(tool, str(tmp_path / "missing.h5"), "frames", flat, 8)
    assert r.returncode != 0 and "Cannot open hdf5 file" in r.stdout
    not_h5 = str(tmp_path / "not.h5")
    open(not_h5, "wb").write(b"x" * 4096)
    assert run(tool, not_h5, "frames", flat, 8).returncode != 0
    assert os.path.getsize(h5) == size0                                         # nothing was written by the failures
    assert run(tool, h5, "frames", flat, 8).returncode == 0
    r = run(tool, h5, "frames", flat, 8)                                        # the name exists now
    assert r.returncode != 0 and "Cannot open hdf5 file" in r.stdout
    f = h5min.File(h5)
    assert f.root["frames"].shape == (50, 8) and not f.root["frames"].data.any()
    # an empty flat file: a dataset without rows and without a chunk tree
    open(str(tmp_path / "empty"), "wb").close()
    assert run(tool, h5, "none", tmp_path / "empty", 8).returncode == 0
    f = h5min.File(h5)
    assert f.root["none"].shape == (0, 8)
    check_root_group(f)
