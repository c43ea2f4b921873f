"""Host-side readers of psp_process's run inputs (upsp-processing_b200/host/upsp_inputs.hpp, run_inputs.hpp):
the input deck (upsp::FileInputs, cpp/lib/upsp_inputs.cpp:35-173, 343-700), the paint calibration and tunnel
conditions (cpp/lib/non_cv_upsp.cpp:19-200), the model-temperature estimate (cpp/exec/psp_process.cpp:2287-2310),
the targets file (cpp/utils/file_readers.ipp:206-255), the plot3d scalar function file (cpp/lib/plot3d.cpp:12-101)
and the first-frame histogram threshold (cpp/lib/image_processing.ipp:10-49, cpp/utils/clustering.ipp:9-101).
Known answers: the deck of the reference's documentation (docs/sphinx/file-formats.rst:251-281) and the layout its
own launcher writes (python/upsp/processing/tree.py:487-528); targets / tunnel values as decoded by the reference's
Python parsers (tests/golden/inputs_golden.json, made by tests/golden/make_inputs_golden.py).  CPU only."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def probe(up):
    return up.build.build_inputs_probe()


def run(probe, *args, ok=True):
    r = subprocess.run([probe] + [str(a) for a in args], capture_output=True, text=True)
    assert (r.returncode == 0) == ok, r.stderr
    return r


def kv(text):
    out = {}
    for line in text.splitlines():
        k, _, v = line.partition(" ")
        out.setdefault(k, []).append(v)
    return out


DOC_DECK = """%Version 1.2.3
%Date_Created: 2026-01-01

@general
    test = my-test-event-name
    run = 1234
    sequence = 56
    tunnel = ames_unitary
@vars
    dir = /nobackup/upsp/test_name
@all
    sds = $dir/inputs/123456.wtd
    grid = $dir/inputs/test-subject.grid
    targets = $dir/inputs/test-subject.tgts
@camera
    number = 2
    filename = $dir/inputs/12345602.mraw
  calibration = $dir/inputs/cam02-to-model.json
    aedc = false
@camera
    number = 1
    cine = $dir/inputs/12345601.cine
    calibration = $dir/inputs/cam01-to-model.json
    aedc = false
# a comment line
@options
    target_patcher = polynomial
    registration = pixel
    overlap = best_view
    filter = gaussian
    filter_size = 3
    oblique_angle = 72.5
    number_frames = 2000
@output
    dir = $dir/outputs
    name = 123456
"""


def test_deck_documented_example(probe, tmp_path):
    (tmp_path / "deck.inp").write_text(DOC_DECK)
    d = kv(run(probe, "deck", tmp_path / "deck.inp").stdout)
    base = "/nobackup/upsp/test_name"
    assert d["version"] == ["1.2.3"] and d["test_id"] == ["my-test-event-name"]
    assert d["run"] == ["1234"] and d["sequence"] == ["56"] and d["tunnel"] == ["ames_unitary"]
    assert d["sds"] == [base + "/inputs/123456.wtd"] and d["grid"] == [base + "/inputs/test-subject.grid"]
    assert d["grid_type"] == ["p3d"] and d["grid_units"] == ["-"] and d["cameras"] == ["2"]
    # cameras sorted by number; @all targets filled into both; legacy `cine` key accepted
    assert d["camera"] == [
        f"1 {base}/inputs/12345601.cine {base}/inputs/test-subject.tgts {base}/inputs/cam01-to-model.json",
        f"2 {base}/inputs/12345602.mraw {base}/inputs/test-subject.tgts {base}/inputs/cam02-to-model.json"]
    assert d["target_patcher"] == ["polynomial"] and d["registration"] == ["pixel"] and d["overlap"] == ["best_view"]
    assert d["pixel_interpolation"] == ["linear"] and d["filter"] == ["gaussian"] and d["filter_size"] == ["3"]
    assert float(d["oblique_angle"][0]) == 72.5 and d["number_frames"] == ["2000"]
    assert d["out_dir"] == [base + "/outputs"] and d["out_name"] == ["123456"]


def test_deck_launcher_layout_and_defaults(probe, tmp_path):
    # python/upsp/processing/tree.py:487-528: tabs, no @vars, unknown keys, empty `normals =` value
    rows = ["%Version 9.9", "%Date_Created: x", "", "@general", "\ttest = t11-0377", "\trun = 4121", "\tsequence = 07",
            "\ttunnel = ames_unitary", "\tframerate = 10000", "\tfstop =  2.8", "@all", "\tgrid = /g/model.tri",
            "\tsds = /g/412107.wtd", "\ttargets = /g/model.tgts", "\tnormals = ", "@camera", "\tnumber = 1",
            "\tcine = /g/41210701.cine", "\tcalibration = /c/cam01.json", "@options", "\ttarget_patcher = none",
            "\tregistration = none", "\tfilter = none", "\tfilter_size = 1", "\toblique_angle = 70", "\tnumber_frames = -1",
            "@output", "\tdir = /o", "\tname = 412107"]
    (tmp_path / "d.inp").write_text("\n".join(rows))
    d = kv(run(probe, "deck", tmp_path / "d.inp").stdout)
    assert d["grid_type"] == ["tri"] and d["normals"] == [""] and d["sequence"] == ["7"] and d["number_frames"] == ["-1"]
    assert d["camera"] == ["1 /g/41210701.cine /g/model.tgts /c/cam01.json"]
    assert d["overlap"] == ["average_views"] and d["registration"] == ["none"] and d["filter"] == ["none"]


@pytest.mark.parametrize("line, msg", [
    ("registration = fancy", "@options:registration"),
    ("filter = median", "@options:filter"),
    ("overlap = all", "@options:overlap"),
    ("filter_size = big", "@options:filter_size"),
    ("target_patcher = spline", "@options:target_patcher"),
])
def test_deck_rejects_bad_options(probe, tmp_path, line, msg):
    (tmp_path / "d.inp").write_text("@general\n test = a\n@options\n " + line + "\n")
    r = run(probe, "deck", tmp_path / "d.inp", ok=False)
    assert msg in r.stderr


def test_deck_unresolved_variable_and_missing_files(probe, tmp_path):
    (tmp_path / "d.inp").write_text("@vars\n a = /x\n@all\n grid = $b/g.tri\n")
    assert "@all:grid" in run(probe, "deck", tmp_path / "d.inp", ok=False).stderr
    assert run(probe, "deck", tmp_path / "nope.inp", ok=False).returncode == 1
    # check_all: every named file must exist (upsp_inputs.cpp:176-228)
    for n in ("g.tri", "s.wtd", "t.tgts", "c.json", "v.mraw"):
        (tmp_path / n).write_text("x")
    deck = (f"@vars\n d = {tmp_path}\n@all\n grid = $d/g.tri\n sds = $d/s.wtd\n targets = $d/t.tgts\n calibration = $d/c.json\n"
            f"@camera\n number = 1\n filename = $d/v.mraw\n@output\n dir = $d\n")
    (tmp_path / "ok.inp").write_text(deck)
    assert kv(run(probe, "deck", tmp_path / "ok.inp", "check").stdout)["grid_type"] == ["tri"]
    (tmp_path / "bad.inp").write_text(deck.replace("s.wtd", "missing.wtd"))
    assert "SDS file" in run(probe, "deck", tmp_path / "bad.inp", "check", ok=False).stderr


def test_paint_calibration_and_gain(probe, tmp_path):
    (tmp_path / "pc.txt").write_text("a = 1.25\nb=-0.003\n c =  2e-6\nd = 0.75\n e = 0.001\nf = -4.5e-7\ncomment line\n")
    T, Pss = 71.3, 11.82
    d = kv(run(probe, "paintcal", tmp_path / "pc.txt", T, Pss).stdout)
    c = {k: np.float32(float(d[k][0])) for k in "abcdef"}
    assert [float(c[k]) for k in "abcdef"] == [float(np.float32(v)) for v in (1.25, -0.003, 2e-6, 0.75, 0.001, -4.5e-7)]
    T, Pss = np.float32(T), np.float32(Pss)
    gain = c["a"] + c["b"] * T + c["c"] * T * T + (c["d"] + c["e"] * T + c["f"] * T * T) * Pss   # f32, reference order
    assert np.float32(float(d["gain"][0])) == gain
    assert run(probe, "paintcal", tmp_path / "absent.txt", ok=False).returncode == 1


def test_tunnel_conditions_match_reference_parser(probe):
    gold = json.load(open(os.path.join(GOLDEN, "inputs_golden.json")))["wtd"]
    d = kv(run(probe, "wtd", os.path.join(GOLDEN, "sample.wtd")).stdout)
    names = dict(ALPHA="alpha", BETA="beta", PHI="phi", PTOT="ptot", TTF="ttot", PS="ps", Q="qbar", RNU="rey", TCAVG="tcavg")
    for k, v in gold.items():
        assert np.float32(float(d[names[k]][0])) == np.float32(v), k
    assert np.float32(float(d["mach"][0])) == np.float32(0.84)
    assert np.float32(float(d["model_temp"][0])) == np.float32(88.125)       # TCAVG supersedes the estimate
    # recovery-factor estimate, psp_process.cpp:2287-2296, float/double mix as written there
    ttot = np.float32(np.float32(97.43) + np.float32(459.67))
    t_inf = np.float32(float(ttot) / (1.0 + (float(np.float32(1.4)) - 1.0) * 0.5 * float(np.float32(0.84)) * float(np.float32(0.84))))
    ttot = np.float32(ttot - np.float32(459.67))
    t_inf = np.float32(t_inf - np.float32(459.67))
    wall = np.float32(np.float32(np.float32(0.896) * np.float32(ttot - t_inf)) + t_inf)
    assert np.float32(float(d["wall_temp"][0])) == wall


def test_tunnel_conditions_reference_fixture(probe):
    path = "/root/reference/test/data/wtd_test.wtd"
    if not os.path.exists(path):
        pytest.skip("reference fixture not present on this machine")
    r = run(probe, "wtd", path)
    d = kv(r.stdout)
    assert float(d["mach"][0]) == 1.0 and np.float32(float(d["alpha"][0])) == np.float32(0.05)
    assert d["tcavg"] == ["nan"] and d["model_temp"] == d["wall_temp"]
    assert "Warning" not in r.stderr            # every condition the reference warns about is present


def test_targets_match_reference_parser(probe):
    gold = json.load(open(os.path.join(GOLDEN, "inputs_golden.json")))["tgts"]
    d = kv(run(probe, "tgts", os.path.join(GOLDEN, "sample.tgts")).stdout)
    assert int(d["count"][0]) == len(gold)
    for line, g in zip(d["target"], gold):
        f = line.split()
        assert int(f[0]) == g["idx"] and [float(v) for v in f[1:4]] == g["xyz"] and float(f[4]) == g["size"]
    fid = kv(run(probe, "tgts", os.path.join(GOLDEN, "sample.tgts"), "*Fiducials").stdout)
    assert fid["count"] == ["2"] and fid["target"][1].split()[:2] == ["2", "0.75"]
    assert kv(run(probe, "tgts", os.path.join(GOLDEN, "sample.tgts"), "*Nothing").stdout)["count"] == ["0"]


def test_targets_reference_fixture(probe):
    path = "/root/reference/test/data/fml_tc3_volume.tgts"
    if not os.path.exists(path):
        pytest.skip("reference fixture not present on this machine")
    d = kv(run(probe, "tgts", path).stdout)
    assert d["count"] == ["24"] and d["target"][0].split()[0] == "1" and d["target"][-1].split()[0] == "24"


def _p3d_function(path, zones, values, seps):
    rec = (lambda b: struct.pack("<i", len(b)) + b + struct.pack("<i", len(b))) if seps else (lambda b: b)
    with open(path, "wb") as f:
        f.write(rec(struct.pack("<i", len(zones))))
        f.write(rec(b"".join(struct.pack("<4i", *z, 1) for z in zones)))
        f.write(rec(np.asarray(values, "<f4").tobytes()))


def test_plot3d_scalar_function_file(probe, tmp_path):
    zones = [(4, 5, 1), (3, 4, 1)]
    vals = (np.arange(32, dtype=np.float32) * 0.25 - 3).astype(np.float32)
    _p3d_function(tmp_path / "raw.f", zones, vals, seps=False)
    out = run(probe, "p3dfun", tmp_path / "raw.f").stdout.split()
    assert out[:2] == ["count", "32"] and np.array_equal(np.array(out[2:], np.float32), vals)
    # with FORTRAN record markers the reference reads the data record through its marker-less overload
    # (cpp/lib/plot3d.cpp:68): slot 0 holds the marker's bits and every scalar sits one slot late.  Same here.
    _p3d_function(tmp_path / "seps.f", zones, vals, seps=True)
    out = np.array(run(probe, "p3dfun", tmp_path / "seps.f").stdout.split()[2:], np.float32)
    assert np.array_equal(out[1:], vals[:-1]) and out[:1].view(np.int32)[0] == 128
    assert "Failed to parse Plot3D function file" in run(probe, "p3dfun", tmp_path / "seps.f", 0, ok=False).stderr
    (tmp_path / "short.f").write_bytes(b"\x01\x00")
    assert run(probe, "p3dfun", tmp_path / "short.f", ok=False).returncode == 1


def _find_peaks(data, separation):
    """cpp/utils/clustering.ipp:9-60, restated"""
    peaks, plateau, begin = [], False, 0
    for i in range(1, len(data) - 1):
        if np.isinf(data[i]) or (data[i] > data[i - 1] and data[i] > data[i + 1]):
            if peaks and i - peaks[-1] < separation:
                if data[peaks[-1]] < data[i]:
                    peaks[-1] = i
                break
            peaks.append(i)
        elif data[i] > data[i - 1] and data[i] == data[i + 1]:
            plateau, begin = True, i
        elif plateau:
            if data[i] < data[i + 1]:
                plateau = False
            elif data[i] > data[i + 1]:
                plateau = False
                pi = (i + begin) // 2
                if peaks and pi - peaks[-1] < separation:
                    if data[peaks[-1]] < data[pi]:
                        peaks[-1] = pi
                    break
                peaks.append(pi)
    return peaks


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_first_frame_threshold(probe, tmp_path, seed):
    rng = np.random.default_rng(seed)
    # dark fiducials + background around 300 + bright paint around 1800-2600 + saturated pixels above the 12-bit range
    img = np.concatenate([rng.normal(300, 40, 3000), rng.normal(1800 + 400 * seed, 250, 60000), rng.normal(60, 10, 500),
                          np.full(7, 5000.0)]).clip(0, 65535).astype(np.uint16)
    img.tofile(tmp_path / "f.u16")
    d = kv(run(probe, "hist", tmp_path / "f.u16", 12).stdout)
    counts = np.bincount(img[img < 4096] // 16, minlength=256)
    assert np.array_equal(np.array(d["counts"][0].split(), int), counts) and d["bin_sz"] == ["16"]
    maxp = _find_peaks(counts.astype(np.int64), 5)
    with np.errstate(divide="ignore"):
        minp = _find_peaks(1.0 / counts, 5)
    first_min = next((m for m in minp if maxp and m > maxp[0]), 0)
    assert [int(v) for v in d["peaks"][0].split()] == maxp
    assert int(d["first_min"][0]) == first_min and int(d["threshold"][0]) == 16 * first_min + 5


@pytest.mark.parametrize("n, maxels", [(5, 1000), (1000, 1000), (2500, 1000), (3999, 1000), (10, 0), (12345, 7)])
def test_regression_sample(probe, tmp_path, n, maxels):
    """the vv-*.dat files (cpp/utils/file_writers.cpp:9-31): step = n // maxels, at most maxels values"""
    v = np.arange(n, dtype=np.float32) * np.float32(0.5)
    v.tofile(tmp_path / "v.f32")
    r = run(probe, "vvdump", tmp_path / "v.f32", tmp_path / "o.dat", maxels)
    numels = maxels if maxels > 0 else n
    step = 1 if n < numels else n // numels
    want = v[::step][:numels]
    assert np.array_equal(np.fromfile(tmp_path / "o.dat", np.float32), want) and r.stdout.split() == ["written", str(len(want))]


def test_deck_write_file_round_trip(probe, tmp_path):
    """FileInputs::write_file (cpp/lib/upsp_inputs.cpp:236-341): the written deck loads back to the same fields;
    variables are put back (longest value first) and shared targets / calibration move to @all."""
    (tmp_path / "deck.inp").write_text(DOC_DECK)
    first = run(probe, "deck", tmp_path / "deck.inp", "write", tmp_path / "again.inp").stdout
    text = (tmp_path / "again.inp").read_text()
    assert "%Version 1.2.3\n%Date_Created 1/2/2026\n" in text
    assert "\tdir = /nobackup/upsp/test_name\n" in text and "\tgrid = $dir/inputs/test-subject.grid\n" in text
    assert text.count("\ttargets = $dir/inputs/test-subject.tgts") == 1 and text.index("targets =") < text.index("@camera")
    assert text.count("\tcalibration = ") == 2                     # per camera: they differ
    again = run(probe, "deck", tmp_path / "again.inp").stdout
    assert kv(again) == kv(first)
