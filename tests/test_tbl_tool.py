"""host/xyz_scalar_to_tbl_b200.cpp against the reference's own tools (cpp/exec/xyz_scalar_to_tbl.cpp,
xyz_scalar_to_tbl_delta.cpp) compiled from the reference tree into oracle/_ref (`make -C oracle ref`): same bytes,
messages and exit codes on the flat files the frame chain writes.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

from test_p3d_model import reference_fixture, write_p3d


@pytest.fixture(scope="module")
def tools(up, orc):
    ref = orc.build_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built and the reference tree is not present on this machine")
    return up.build.build_tbl_tool(), ref


def flat_files(d, n, seed):
    rng = np.random.default_rng(seed)
    xyz = rng.normal(0, 20, (n, 3)).astype(np.float32)
    s1 = rng.normal(0, 1.5, n).astype(np.float32)
    s2 = rng.normal(0, 1e-3, n).astype(np.float32)
    s1[[1, n // 2]] = np.nan                     # uncovered nodes are NaN in the outputs
    s2[[1, n - 1]] = np.nan
    s1[3], s2[4] = np.float32(1e20), np.float32(-1e-20)
    for name, v in (("X", xyz[:, 0]), ("Y", xyz[:, 1]), ("Z", xyz[:, 2]), ("rms", s1), ("steady_state", s2)):
        v.tofile(d / name)


def test_table_matches_reference_tool(tools, tmp_path):
    mine, ref = tools
    write_p3d(tmp_path / "g.p3d", reference_fixture())
    flat_files(tmp_path, 52, 1)
    args = [str(tmp_path / n) for n in ("g.p3d", "X", "Y", "Z", "rms")]
    r1 = subprocess.run([os.path.join(ref, "xyz_scalar_to_tbl")] + args + [str(tmp_path / "ref.tecplot")], capture_output=True, text=True)
    r2 = subprocess.run([mine] + args + [str(tmp_path / "mine.tecplot")], capture_output=True, text=True)
    assert r1.returncode == r2.returncode == 0
    a, b = (tmp_path / "ref.tecplot").read_bytes(), (tmp_path / "mine.tecplot").read_bytes()
    assert a == b and a.count(b"ZONE T=") == 3 and len(a.splitlines()) == 2 + 3 + 52


def test_delta_table_matches_reference_tool(tools, tmp_path):
    mine, ref = tools
    write_p3d(tmp_path / "g.p3d", reference_fixture())
    flat_files(tmp_path, 52, 2)
    args = [str(tmp_path / n) for n in ("g.p3d", "X", "Y", "Z", "rms", "steady_state")]
    for exe, sub in (([os.path.join(ref, "xyz_scalar_to_tbl_delta")], "r"), ([mine, "-delta"], "m")):
        (tmp_path / sub).mkdir()
        assert subprocess.run(exe + args, cwd=tmp_path / sub, capture_output=True).returncode == 0
    assert (tmp_path / "r" / "xyz_scalar_delta.tecplot").read_bytes() == (tmp_path / "m" / "xyz_scalar_delta.tecplot").read_bytes()


def test_error_behaviour_matches_reference_tool(tools, tmp_path):
    mine, ref = tools
    write_p3d(tmp_path / "g.p3d", reference_fixture())
    flat_files(tmp_path, 52, 3)
    np.zeros(51, np.float32).tofile(tmp_path / "short")
    cases = [["g.p3d", "X", "Y", "Z", "short", "o"], ["g.p3d", "X", "short", "Z", "rms", "o"], ["nope.p3d", "X", "Y", "Z", "rms", "o"],
             ["g.p3d", "X", "Y", "nope", "rms", "o"], ["g.p3d", "X"]]
    for c in cases:
        args = [str(tmp_path / n) for n in c]
        r1 = subprocess.run([os.path.join(ref, "xyz_scalar_to_tbl")] + args, capture_output=True, text=True, cwd=tmp_path)
        r2 = subprocess.run([mine] + args, capture_output=True, text=True, cwd=tmp_path)
        assert r1.returncode == r2.returncode == 1
        assert r1.stdout == r2.stdout, c


@pytest.mark.parametrize("flag, msize, nframes", [(0, 1237, 301), (1, 1237, 301), (0, 100, 100), (1, 7, 2500)])
def test_reference_transpose_tool_pins_the_restatement(tools, orc, tmp_path, flag, msize, nframes):
    """The reference's own local_transpose / global_transpose (cpp/exec/upsp_matrix_transpose.cpp:86-212, the code of
    psp_process.cpp:647-771) compiled from the reference tree against a single-rank loop-back MPI header
    (oracle/mpi_stub/mpi.h) and run here: its `pressure_transpose` / `pressure` file is what the CPU restatement
    (oracle.global_transpose) and numpy's transpose give, bit for bit.  The GPU tool is held against the same arrays in
    tests/test_host_driver.py::test_transpose_tool_matches_numpy."""
    _, ref = tools
    exe = os.path.join(ref, "upsp_matrix_transpose")
    if not os.path.exists(exe):
        pytest.skip("reference transpose tool not built")
    rng = np.random.default_rng(msize + flag)
    a = rng.standard_normal((nframes, msize) if flag == 0 else (msize, nframes)).astype(np.float32)
    a[0, 0] = np.nan
    a.tofile(tmp_path / "in")
    r = subprocess.run([exe, str(msize), str(nframes), str(flag), str(tmp_path / "in"), str(tmp_path)], capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(tmp_path / ("pressure_transpose" if flag == 0 else "pressure"), np.float32).reshape(a.T.shape)
    assert np.array_equal(out.view(np.uint32), np.ascontiguousarray(a.T).view(np.uint32))
    if flag == 0:
        mine = np.concatenate(orc.global_transpose([a], msize, nframes), 0)
        assert np.array_equal(mine.view(np.uint32), out.view(np.uint32))
