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


def _write_cine(path, codes, bpp, width, height):
    """Minimal Vision Research cine (CINEFILEHEADER 44 B, BITMAPINFOHEADER 40 B, SETUP 7240 B, offset
    table, {annotation size, image size, image} per frame) holding packed 10- or 12-bit `codes`."""
    import struct
    nf = codes.shape[0]
    frames = []
    for f in range(nf):
        v = codes[f].reshape(-1).astype(np.uint32)
        if bpp == 12:
            a, b = v[0::2], v[1::2]
            by = np.stack([a >> 4, ((a & 0xF) << 4) | (b >> 8), b & 0xFF], 1)
        else:
            a, b, c, d = v[0::4], v[1::4], v[2::4], v[3::4]
            by = np.stack([a >> 2, ((a & 3) << 6) | (b >> 4), ((b & 0xF) << 4) | (c >> 6),
                           ((c & 0x3F) << 2) | (d >> 8), d & 0xFF], 1)
        frames.append(by.astype(np.uint8).tobytes())
    setup = bytearray(7240)
    struct.pack_into("<H", setup, 0, 1000)        # FrameRate16
    struct.pack_into("<H", setup, 140, 0x5453)    # Mark
    struct.pack_into("<H", setup, 142, 7240)      # Length
    struct.pack_into("<HH", setup, 737, width, height)
    struct.pack_into("<I", setup, 896, bpp)       # RealBPP
    off_offsets = 44 + 40 + 7240
    cfh = struct.pack("<HHHHiIiIIII", 0x4943, 44, 0, 1, 0, nf, 0, nf, 44, 84, off_offsets) + bytes(8)
    bmi = struct.pack("<IiiHHIIiiII", 40, width, height, 1, 16, 256, len(frames[0]), 0, 0, 0, 0)
    first = off_offsets + 8 * nf
    offs = [first + i * (8 + len(frames[0])) for i in range(nf)]
    with open(path, "wb") as fd:
        fd.write(cfh + bmi + bytes(setup) + struct.pack("<%dq" % nf, *offs))
        for fr in frames:
            fd.write(struct.pack("<II", 8, len(fr)) + fr)


@pytest.mark.gpu
@pytest.mark.parametrize("container", ["mraw", "cine10"])
def test_host_driver_reads_camera_files(up, orc, gpu, tmp_path, container):
    """job.txt `video0 = file`: the driver reads a Photron .mraw / Vision Research .cine with
    host/video_readers.hpp and pushes the stored bytes; decode, the 10->12-bit table and everything
    after it run on the GPU.  Results must equal the oracle run on the decoded frames."""
    import upsp_b200
    synth = upsp_b200.synth
    case = Case(synth, n_frames=24, n_nodes=1500, registration=True, patches=True, seed=43)
    lut = np.load(os.path.join(os.path.dirname(__file__), "golden", "video_golden.npz"))["lut10"]
    job, out = tmp_path / "job", tmp_path / "out"
    out.mkdir()
    if container == "cine10":      # frames become table[codes]: the oracle sees what the reference would decode
        rng = np.random.default_rng(9)
        codes = np.clip(case.frames[0].astype(np.int64) // 4 + rng.integers(-2, 3, case.frames[0].shape), 0, 1023)
        case.frames = [lut[codes].astype(np.uint16)]
    synth.write_job(str(job), frames=case.frames, csr=case.csr, fmt="p12", registration="given",
                    warp=case.warp, patches=[synth.flatten_patches(*p) for p in case.patch_lists], remap=None,
                    cal=case.cal, qbar=case.qbar, ps=case.ps, steady=case.steady, model_temp=case.temp)
    os.remove(job / "cam0.frames")
    if container == "mraw":
        synth.pack_12bit(case.frames[0].reshape(case.F, -1)).tofile(job / "cam0.mraw")
        (job / "cam0.cih").write_text("#Camera Information Header\r\nRecord Rate(fps) : 1000\r\nTotal Frame : %d\r\n"
                                      "Image Width : %d\r\nImage Height : %d\r\nColor Bit : 12\r\n" % (case.F, case.W, case.H))
        name = "cam0.mraw"
    else:
        _write_cine(job / "cam0.cine", codes, 10, case.W, case.H)
        name = "cam0.cine"
    with open(job / "job.txt", "a") as fd:
        fd.write("video0 = %s\n" % name)
    exe = up.build.build_host()
    r = subprocess.run([exe, "-job_dir", str(job), "-out_dir", str(out), "-chunk", "8"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    ref = run_oracle(orc, case)
    rd = lambda n, shape=None: (np.fromfile(out / n, np.float32).reshape(shape) if shape
                                else np.fromfile(out / n, np.float32))
    assert same_bits(rd("intensity_transpose", (case.N, case.F)), ref["itrans"])
    assert same_bits(rd("intensity_avg"), ref["avg"]) and same_bits(rd("intensity_rms"), ref["rms"])
    assert same_bits(rd("gain"), ref["gain"])


def test_transpose_tool_fails_loudly(up, tmp_path):
    exe = up.build.build_transpose_tool()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
    r = subprocess.run([exe, "5", "7", "0", str(tmp_path / "nope"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot open" in r.stderr
    (tmp_path / "short").write_bytes(bytes(12))
    r = subprocess.run([exe, "5", "7", "0", str(tmp_path / "short"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "expected msize*number_frames*4" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("flag,block", [(0, None), (1, None), (0, "4000"), (1, "3000")])
def test_transpose_tool_matches_numpy(up, gpu, tmp_path, flag, block):
    """upsp_matrix_transpose (cpp/exec/upsp_matrix_transpose.cpp): pressure [F x N] <-> pressure_transpose
    [N x F], whole matrix at once and in several row blocks; ragged sizes as in test_general_utils.cpp."""
    msize, nframes = 1237, 301
    rng = np.random.default_rng(3)
    a = rng.standard_normal((nframes, msize) if flag == 0 else (msize, nframes)).astype(np.float32)
    a.tofile(tmp_path / "in")
    exe = up.build.build_transpose_tool()
    env = dict(os.environ)
    if block:
        env["UPSP_XPOSE_BLOCK_FLOATS"] = block
    r = subprocess.run([exe, str(msize), str(nframes), str(flag), str(tmp_path / "in"), str(tmp_path)],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(tmp_path / ("pressure_transpose" if flag == 0 else "pressure"), np.float32)
    assert np.array_equal(out.reshape(a.T.shape), a.T)


@pytest.mark.gpu
def test_host_driver_builds_patches_from_targets(up, orc, gpu, tmp_path):
    """job directory with cam0.targets (projected fiducials u v diameter) instead of precomputed
    pixel lists: the driver clusters them and builds the boundary / interior lists with
    host/patch_geometry.hpp (InitializeImagePatches, psp_process.cpp:2125-2163); the run must equal
    the oracle run with the lists of the oracle's own restatement of that geometry."""
    import upsp_b200
    from oracle import setup_patches as sp
    synth = upsp_b200.synth
    case = Case(synth, n_frames=24, n_nodes=1500, registration=True, patches=False, seed=47)
    rng = np.random.default_rng(5)
    targs = [(rng.uniform(12, case.W - 12), rng.uniform(12, case.H - 12), rng.uniform(3, 6)) for _ in range(6)]
    targs += [(40.0, 30.0, 5.0), (46.0, 31.0, 4.0), (2.0, 50.0, 5.0)]           # a two-target cluster and a border target
    targs = [tuple(np.float32(x) for x in t) for t in targs]
    bt, bf = 2, 1
    first = case.frames[0][0]
    thresh = int(np.percentile(first, 2))
    lists = sp.patch_clusters(sp.cluster_points(targs, bt + bf), case.W, case.H, bt, bf, first, thresh, 2)
    xy = lambda pts: (np.array([p[0] for p in pts], np.uint32), np.array([p[1] for p in pts], np.uint32))
    case.patch_lists = [([xy(b) for b, _ in lists], [xy(i) for _, i in lists])]
    job, out = tmp_path / "job", tmp_path / "out"
    out.mkdir()
    synth.write_job(str(job), frames=case.frames, csr=case.csr, fmt="p12", registration="given", warp=case.warp,
                    patches=[synth.flatten_patches(*case.patch_lists[0])], remap=None, cal=case.cal, qbar=case.qbar,
                    ps=case.ps, steady=case.steady, model_temp=case.temp)
    for f in os.listdir(job):
        if ".patch_" in f:
            os.remove(job / f)
    (job / "cam0.targets").write_text("".join("%.9g %.9g %.9g\n" % tuple(float(x) for x in t) for t in targs))
    first.astype("<u2").tofile(job / "cam0.first")
    with open(job / "job.txt", "a") as fd:
        fd.write("bound_thickness = %d\nbuffer_thickness = %d\npatch_thresh = %d\n" % (bt, bf, thresh))
    exe = up.build.build_host()
    r = subprocess.run([exe, "-job_dir", str(job), "-out_dir", str(out), "-chunk", "8"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Sorted 9 targets into 8 clusters" in r.stdout
    ref = run_oracle(orc, case)
    got = np.fromfile(out / "intensity_transpose", np.float32).reshape(case.N, case.F)
    assert same_bits(got, ref["itrans"])
    assert same_bits(np.fromfile(out / "intensity_avg", np.float32), ref["avg"])
