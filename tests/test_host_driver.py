"""The C++ host driver (upsp-processing_b200/host): builds with g++ against the C ABI only;
on a GPU it runs a whole synthetic psp_process job from a job directory and writes the
reference's flat files, which must equal the oracle's results."""
import os
import subprocess

import numpy as np
import pytest

from chain import Case, cp_errors, monomial_mass, run_oracle, same_bits


def test_host_driver_builds_and_fails_loudly_without_inputs(up, tmp_path):
    exe = up.build.build_host()
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
    r = subprocess.run([exe, "-job_dir", str(tmp_path / "nope"), "-out_dir", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot open" in r.stderr     # psp_process: return 1 on bad inputs


@pytest.mark.gpu
@pytest.mark.parametrize("registration", ["given", "pixel"])
def test_host_driver_writes_reference_flat_files(up, orc, gpu, tmp_path, registration):
    import upsp_b200
    synth = upsp_b200.synth
    case = Case(synth, n_frames=40, n_nodes=2500, registration=True, patches=True, overlap=True, seed=41,
                texture=600.0)
    job, out = tmp_path / "job", tmp_path / "out"
    out.mkdir()
    remap = orc.overlap_remap(case.N, case.overlap)
    synth.write_job(str(job), frames=case.frames, csr=case.csr, fmt="p12", registration=registration,
                    warp=case.warp, patches=[synth.flatten_patches(*p) for p in case.patch_lists], remap=remap,
                    cal=case.cal, qbar=case.qbar, ps=case.ps, steady=case.steady, model_temp=case.temp)
    exe = up.build.build_host()
    r = subprocess.run([exe, "-job_dir", str(job), "-out_dir", str(out), "-chunk", "16"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "## 'pressure_transpose' written" in r.stderr
    rd = lambda n, shape=None: (np.fromfile(out / n, np.float32).reshape(shape) if shape
                                else np.fromfile(out / n, np.float32))
    got = dict(intensity=None, avg=rd("intensity_avg"), rms=rd("intensity_rms"), coverage=rd("coverage"),
               itrans=rd("intensity_transpose", (case.N, case.F)), ptrans=rd("pressure_transpose", (case.N, case.F)),
               rms2=rd("rms"), avg2=rd("avg"), gain=rd("gain"))
    if registration == "pixel":
        # hand the maps the driver's on-device ECC found to the oracle: re-run the registration
        # through the Python binding (same library, same inputs -> same maps)
        g = up.PspGpu(1, 1, case.F)
        g.set_camera(0, case.W, case.H)
        g.set_projection(0, np.array([0, 1], np.int32), np.array([0], np.int32), np.ones(1, np.float32))
        g.set_options(registration=up.REG_PIXEL)
        g.set_reference_frame(0, case.frames[0][0])
        g.push_frames(0, case.frames[0], up.PIX_U16, 0, case.F)
        g.process_frames(0, case.F)
        case.warp = [g.read_warp_matrices(0)]
        g.close()
    ref = run_oracle(orc, case)
    for k in ("avg", "rms", "coverage", "itrans", "gain"):
        assert same_bits(got[k], ref[k]), k
    exact = run_oracle(orc, case, exact_fit=True)
    e_exact, _ = cp_errors(case, exact, got)
    cond = 8 * np.finfo(np.float32).eps * monomial_mass(orc, ref)
    assert np.all(e_exact <= 1e-6 + cond)
    st = rd("steady_state")
    assert np.array_equal(np.isnan(st), case.steady > 3.0) and same_bits(rd("model_temp"), case.temp)
