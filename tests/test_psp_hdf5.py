"""HDF5 outputs of psp_process (-h5_out, extras.h5) written without an HDF5 library (host/psp_hdf5.hpp).

Reference: upsp::PSPWriter cpp/lib/PSPHDF5.ipp:14-78, 104-312, 361-411, 447-781 and its use in
cpp/exec/psp_process.cpp:2400-2420, 2535-2604.  tests/h5min.py is a minimal reader of the same subset of the file
format; it is first held against the reference's own fixtures (files written by the real HDF5 library), then used to read
what host/psp_hdf5.hpp writes, and the two kinds of file are compared message by message where the reference writes the
same thing (datatype / dataspace / layout messages, attribute encoding, group structure)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import h5min

REF_INPUTS = "/root/reference/cpp/test/inputs"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_INPUTS), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def probe():
    import upsp_b200
    return upsp_b200.build.build_h5_probe()


@needs_ref
@pytest.mark.parametrize("name,transposed", [("unstruct_nodal_pencil.h5", 0), ("unstruct_nodal_pencil_trans.h5", 1)])
def test_reader_parses_the_reference_fixtures(name, transposed):
    f = h5min.File(os.path.join(REF_INPUTS, name))
    assert f.root.attrs["psph5_version"][0] == 1 and f.root.attrs["nodal"][0] == 1
    assert f.root.attrs["transpose"][0] == transposed and f.root.attrs["structured"][0] == 0
    assert f.root.attrs["code_version"] == ["sample scripts"]
    assert f.root["Grid"].attrs["units"] == ["in"]
    x = f.root["Grid/x"].data
    tri = f.root["Grid/triangles"].data
    assert x.shape == (1155,) and tri.shape == (2080, 3) and tri.max() < 1155
    assert f.root["Condition/mach"].data[0] == np.float32(0.85) and f.root["Condition/mach"].attrs["units"] == ["-"]
    assert f.root["Condition/test_id"].data[0].split(b"\0")[0] == b"sample file"
    fr = f.root["frames"].data                       # chunked + deflate in the fixtures
    assert fr.shape == ((1155, 200) if transposed else (200, 1155))
    assert np.isfinite(fr).all() and fr.max() > 0


def _objects(f):
    return {p: o for p, o in h5min.walk(f.root)}


@pytest.mark.parametrize("kind,n", [("unstructured", 50), ("structured", 40), ("unstructured", 100003)])
def test_writer_round_trip(probe, tmp_path, kind, n):
    out = str(tmp_path / "t.h5")
    subprocess.run([probe, out, kind, str(n)], check=True)
    f = h5min.File(out)
    assert f.eof == os.path.getsize(out)
    assert f.root.attrs["psph5_version"][0] == 1 and f.root.attrs["nodal"][0] == 1 and f.root.attrs["transpose"][0] == 1
    assert f.root.attrs["structured"][0] == (1 if kind == "structured" else 0)
    assert f.root.attrs["code_version"] == ["probe 1.0"]
    i = np.arange(n)
    np.testing.assert_array_equal(f.root["Grid/x"].data, (0.5 * i).astype(np.float32))
    np.testing.assert_array_equal(f.root["Grid/y"].data, (1.0 - i).astype(np.float32))
    np.testing.assert_array_equal(f.root["rms"].data, (np.float32(0.001) * i.astype(np.float32)))
    np.testing.assert_array_equal(f.root["coverage"].data, (i % 3).astype(np.float32))
    assert f.root["rms"].attrs["units"] == ["delta Cp"] and f.root["steady_state"].attrs["units"] == ["Cp"]
    assert f.root["model_temp"].attrs["units"] == ["F"] and "units" not in f.root["coverage"].attrs
    assert f.root["Grid"].attrs["units"] == ["in"]
    if kind == "structured":
        np.testing.assert_array_equal(f.root["Grid/grid_sizes"].data, [[n // 2, 2, 1]])
    else:
        t = np.arange(n - 2)
        np.testing.assert_array_equal(f.root["Grid/triangles"].data, np.stack([t, t + 1, t + 2], 1).astype(np.uint32))
        np.testing.assert_array_equal(f.root["Grid/components"].data, (t % 4).astype(np.int32))
    c = f.root["Condition"]
    assert c["test_id"].data[0].split(b"\0")[0] == b"t11-0377"
    assert c["run"].data[0] == 12 and c["sequence"].data[0] == 3 and c["frame_rate"].data[0] == 10000
    assert c["thermocouple_average_temperature"].data[0] == np.float32(71.5)
    np.testing.assert_array_equal(c["focal_length"].data, np.array([-3512.25, -3498.5], np.float32))
    units = {"alpha": "deg", "beta": "deg", "phi": "deg", "mach": "-", "reynolds_number": "millions/ft", "total_pressure": "psf",
             "dynamic_pressure": "psf", "total_temperature": "degF", "thermocouple_average_temperature": "degF",
             "static_pressure": "psf", "run": "-", "sequence": "-", "frame_rate": "Hz", "fstop": "-", "exposure": "microseconds",
             "focal_length": "mm"}
    for k, u in units.items():
        assert c[k].attrs["units"] == [u], k


@needs_ref
def test_same_encoding_as_the_hdf5_library(probe, tmp_path):
    """Datasets and attributes the reference's fixture and the probe's file have in common are encoded with identical
    datatype / dataspace messages and attribute bodies (apart from the values), and every object of the fixture except the
    'frames' histories (which psp_process no longer writes, psp_process.cpp:1881-1923) has its counterpart."""
    out = str(tmp_path / "t.h5")
    subprocess.run([probe, out, "unstructured", "1155"], check=True)
    mine, ref = h5min.File(out), h5min.File(os.path.join(REF_INPUTS, "unstruct_nodal_pencil.h5"))
    om, orf = _objects(mine), _objects(ref)
    assert set(orf) - {"/frames"} <= set(om)

    def msgs(f, path, types):
        addr = f.addr_of(path)
        return [(t, d) for t, d in f._messages(addr) if t in types]

    for path in sorted(set(orf) - {"/frames"}):
        if isinstance(orf[path], h5min.Group):
            continue
        same_shape = path not in ("/Condition/focal_length", "/Grid/triangles", "/Grid/components")   # probe: 2 cameras, n - 2 faces
        assert om[path].dtype == orf[path].dtype and len(om[path].shape) == len(orf[path].shape), path
        assert msgs(mine, path, (0x3,)) == msgs(ref, path, (0x3,)), f"{path}: datatype message differs from the library's"
        sa, sb = msgs(mine, path, (0x1,))[0][1], msgs(ref, path, (0x1,))[0][1]
        assert sa[:8] == sb[:8] and len(sa) == len(sb), f"{path}: dataspace message differs from the library's"
        if not same_shape:
            continue
        assert om[path].shape == orf[path].shape and sa == sb, path
        la, lb = msgs(mine, path, (0x8,))[0][1], msgs(ref, path, (0x8,))[0][1]
        assert la[:2] == lb[:2] and struct.unpack_from("<Q", la, 10) == struct.unpack_from("<Q", lb, 10), path   # version, class, size
        ua = [d for t, d in msgs(mine, path, (0xC,))]
        ub = [d for t, d in msgs(ref, path, (0xC,))]
        assert len(ua) == len(ub), path
        for x, y in zip(ua, ub):
            assert x[:48] == y[:48], f"{path}: attribute header / name / datatype / dataspace differ"
