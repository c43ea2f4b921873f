"""host/video_readers.hpp (cine / mraw containers, SURVEY 8f "I/O edges") against the reference's
readers.  The fixtures under tests/golden/ were written with the reference's own header structures
and decoded with its Python readers (tests/golden/make_golden.py video); the C++ readers must report
the same properties and hand over exactly the stored bytes, which the oracle's unpackers (pinned
separately against the reference) turn into the same pixels.  CPU only."""
import os
import subprocess
import zlib

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def probe():
    import upsp_b200
    return upsp_b200.build.build_probe()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "video_golden.npz"))


def run_probe(probe, path, *args):
    r = subprocess.run([probe, path, *map(str, args)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    kv, crc = {}, {}
    for line in r.stdout.splitlines():
        t = line.split()
        if t[0] == "crc":
            crc[int(t[1])] = int(t[2])
        else:
            kv[t[0]] = float(t[1])
    return kv, crc


@pytest.mark.parametrize("bpp", [10, 12])
def test_cine_reader_matches_reference(probe, gold, orc, tmp_path, bpp):
    path = os.path.join(GOLD, f"synthetic{bpp}.cine")
    dump = tmp_path / "frames.bin"
    kv, crc = run_probe(probe, path, 1, 3, dump)
    w, h, depth, nf, rate = gold[f"cine{bpp}_props"]
    assert (kv["width"], kv["height"], kv["bit_depth"], kv["num_frames"], kv["frame_rate"]) == (w, h, depth, nf, rate)
    assert kv["aperture"] == pytest.approx(2.8) and kv["exposure"] == pytest.approx(25.0, rel=1e-6)
    assert kv["pixel_format"] == (2 if bpp == 10 else 1) and kv["has_lut"] == (1 if bpp == 10 else 0)
    assert [crc[i + 1] for i in range(3)] == [int(c) for c in gold[f"cine{bpp}_crc"]]
    raw = np.fromfile(dump, np.uint8).reshape(3, -1)
    for f in range(3):          # stored bytes -> pixels, as the GPU decoder (== oracle) does
        px = orc.unpack_10bit(raw[f], gold["lut10"]) if bpp == 10 else orc.unpack_12bit(raw[f])
        assert np.array_equal(px.reshape(int(h), int(w)), gold[f"cine{bpp}_frames"][f])


def test_cine_lut_is_the_reference_table(probe, gold):
    kv, _ = run_probe(probe, os.path.join(GOLD, "synthetic10.cine"), 1, 1)
    r = subprocess.run([probe, os.path.join(GOLD, "synthetic10.cine"), "1", "1"], capture_output=True, text=True)
    lut_crc = int([l for l in r.stdout.splitlines() if l.startswith("lut_crc")][0].split()[1])
    assert lut_crc == zlib.crc32(gold["lut10"].astype("<u2").tobytes())


def test_mraw_reader_matches_reference(probe, gold, orc, tmp_path):
    dump = tmp_path / "frames.bin"
    kv, crc = run_probe(probe, os.path.join(GOLD, "tiny12.mraw"), 1, 2, dump)
    w, h, depth, nf, rate = gold["mraw_props"]
    assert (kv["width"], kv["height"], kv["bit_depth"], kv["num_frames"], kv["frame_rate"]) == (w, h, depth, nf, rate)
    assert [crc[1], crc[2]] == [int(c) for c in gold["mraw_crc"]]
    raw = np.fromfile(dump, np.uint8).reshape(2, -1)
    for f in range(2):
        assert np.array_equal(orc.unpack_12bit(raw[f]).reshape(int(h), int(w)), gold["mraw_frames"][f])
    # second frame only, 1-based like VideoReader::read_frame
    _, crc2 = run_probe(probe, os.path.join(GOLD, "tiny12.mraw"), 2, 1)
    assert crc2 == {2: int(gold["mraw_crc"][1])}


def test_reference_mraw_fixture(probe):
    """cpp/test/mraw/12bitMRAW.{mraw,cih} (the reference's own fixture; build container only)."""
    path = "/root/reference/cpp/test/mraw/12bitMRAW.mraw"
    if not os.path.exists(path):
        pytest.skip("reference fixture not present on this machine")
    g = np.load(os.path.join(GOLD, "mraw_golden.npz"))
    kv, crc = run_probe(probe, path)
    assert (kv["width"], kv["height"], kv["bit_depth"], kv["num_frames"]) == (1024, 1024, 12, 2)
    assert [crc[1], crc[2]] == [int(c) for c in g["crc_packed"]]


def test_reader_errors(probe, tmp_path):
    r = subprocess.run([probe, str(tmp_path / "missing.cine")], capture_output=True, text=True)
    assert r.returncode == 1 and "Video File is invalid" in r.stderr       # CineReader.cpp:92-94
    r = subprocess.run([probe, str(tmp_path / "movie.avi")], capture_output=True, text=True)
    assert r.returncode == 1 and "unknown video type" in r.stderr          # psp_process.cpp:429-432
    bad = tmp_path / "bad.cine"
    bad.write_bytes(b"XX" + bytes(8000))
    r = subprocess.run([probe, str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "magic" in r.stderr
    r = subprocess.run([probe, os.path.join(GOLD, "tiny12.mraw"), "3", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "outside" in r.stderr
