"""Command-line validation of psp_process_b200 before anything touches the GPU: ParseOpts' order and messages
(cpp/exec/psp_process.cpp:1192-1249: help, then -input_file / -h5_out / -paint_cal, then the deck) and main()'s exit status 1
with "Failed to validate command line options" (:1362-1367)."""
import subprocess

import pytest


@pytest.fixture(scope="module")
def exe():
    import upsp_b200
    return upsp_b200.build.build_host()


def run(exe, *args):
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=60)


@pytest.mark.parametrize("flag", ["-help", "-h", "-usage", "--help"])
def test_help_prints_the_option_list_and_fails_like_the_reference(exe, flag):
    r = run(exe, flag)
    assert r.returncode == 1
    for key in ("input_file", "frames", "code_version", "trans_nodes", "add_out_dir", "checkout", "bound_pts", "buffer_pts",
                "target_diam_sf", "cutoff_x_max", "h5_out", "steady_p3d", "steady_grid", "paint_cal", "model_temp_p3d"):
        assert "-" + key in r.stdout
    assert "Failed to validate command line options" in r.stderr


def test_required_options_are_reported_in_the_reference_order(exe, tmp_path):
    deck = str(tmp_path / "missing.inp")
    r = run(exe, f"-input_file={deck}")
    assert r.returncode == 1 and "Must specify -h5_out" in r.stderr
    r = run(exe, f"-input_file={deck}", f"-h5_out={tmp_path / 'o.h5'}")
    assert r.returncode == 1 and "Must specify -paint_cal" in r.stderr
    r = run(exe, f"-input_file={deck}", f"-h5_out={tmp_path / 'o.h5'}", f"-paint_cal={tmp_path / 'p.cal'}")
    assert r.returncode == 1 and "cannot be opened" in r.stderr          # FileInputs::Load (upsp_inputs.cpp:35-60)
    r = run(exe)
    assert r.returncode == 1 and "usage" in r.stderr
    r = run(exe, "-input_file")                                            # a key without its value
    assert r.returncode == 1 and "missing value" in r.stderr
