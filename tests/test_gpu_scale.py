"""Parity at BASELINE.json scale and on the real multi-process path.

* test_baseline_scale_parity: configs[1]'s geometry (1024 x 1024 12-bit packed frames, 500 000-node
  grid, batch 256, registration, fiducial patches) on 512 frames, EVERY node and frame against the
  oracle: intensity_transpose / avg / rms / gain bit-exact, delta-Cp by the north-star metric.  This is the
  configuration whose 32-bit offsets, raster permutation, TMA boxes and corner logic matter.
* test_torchrun_two_processes: `python -m torch.distributed.run --nproc-per-node 2 bench.py --check`:
  two PROCESSES, two devices, buffers wired through upsp_gpu_ipc_export / upsp_gpu_ipc_import (pidfd +
  cuMemImportFromShareableHandle), rows stored into the peer over NVLink; every rank checks what it holds
  against the oracle (bench.py parity_check).  Skipped with fewer than 2 GPUs.
Reference: cpp/exec/psp_process.cpp:707-771 (global_transpose), 1866-1872 (reduce), 2460-2498 (phase 2)."""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

from chain import same_bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_baseline_scale_parity(up, orc, gpu):
    import bench
    from upsp_b200 import synth
    F = 512
    args = types.SimpleNamespace(height=1024, width=1024, nodes=500_000, frames=F, targets=32, distinct=128,
                                 degree=6, batch=0, csr="surface", registration="given", cams=1, overlap_groups=0, config=1, exchange="peer")
    wl = bench.build_workload(args, synth)
    # ---- oracle: the whole job
    orc.set_num_threads(bench.host_threads())
    idx = np.arange(F) % args.distinct
    fr = orc.unpack_12bit_frames(wl["packed"][0][idx]).reshape(F, args.height, args.width)
    warps = synth.make_warps(F, seed=5)
    inten, s, q = orc.phase1([fr], wl["csr"], warp=[warps], interp=1, patches=bench.oracle_patches(orc, wl, args))
    del fr
    avg, rms = orc.phase1_finals(s, q, F)
    cov = orc.coverage(wl["csr"])
    itr = orc.global_transpose([inten], args.nodes, F)[0]
    del inten
    # ---- product, through the C ABI
    g = bench.configure(up, wl, args, 0, 1, 0, 0, None)
    for o in range(0, F, args.distinct):
        g.push_frames(0, wl["packed"][0][:min(args.distinct, F - o)], up.PIX_PACKED12, o, min(args.distinct, F - o))
    g.process_frames(0, F)
    g.finish_phase1()
    g.transpose()
    a_g, r_g, c_g = g.read_phase1_stats()
    assert same_bits(a_g, avg) and same_bits(r_g, rms) and same_bits(c_g, cov)
    got = g.read_intensity_transpose()
    assert same_bits(got, itr), "intensity_transpose differs at BASELINE scale"
    del got
    g.phase2(wl["cal"], wl["qbar"], wl["ps"], wl["steady"], wl["temp"], args.degree)
    pt = g.read_pressure_transpose()
    rms2, avg2, gain = g.read_phase2_stats()
    g.close()
    # phase 2 of a spread of 20 000 nodes through the oracle (the per-node float QR is the slow part), in the
    # reference's float arithmetic and with the float64 least-squares fit (DESIGN.md section 4 "Tolerances")
    sel = np.linspace(0, args.nodes - 1, 20_000).astype(np.int64)
    oargs = (itr[sel], avg[sel], cov[sel], wl["steady"][sel], wl["temp"][sel], wl["cal"], wl["qbar"], wl["ps"], args.degree)
    p_ref, _, _, gain_ref = orc.phase2(*oargs)
    p_exact, _, _, _ = orc.phase2(*oargs, exact_fit=True)
    assert same_bits(gain[sel], gain_ref)
    valid = (cov[sel] != 0) & np.all(np.isfinite(itr[sel]) & (itr[sel] != 0), axis=1)
    assert valid.sum() > 15_000
    Kn = np.abs(gain_ref[valid]).astype(np.float64) * 144.0 / float(wl["qbar"])
    r = (avg[sel][valid, None] / itr[sel][valid]).astype(np.float32)
    scale = Kn * np.abs(r).max(axis=1)
    err = lambda a, b: np.abs(a[valid] - b[valid]).max(axis=1)
    e_exact = err(pt[sel], p_exact) / scale                # product vs float64 model
    e_ref = err(pt[sel], p_ref) / scale                    # product vs the reference's float QR (restated)
    noise = err(p_ref, p_exact) / scale                    # that float QR vs the float64 model
    cpmax = np.abs(p_ref[valid]).max(axis=1)
    ns_gpu = err(pt[sel], p_ref) / cpmax                   # north-star metric as SURVEY section 7 states it
    ns_orc = err(p_ref, p_exact) / cpmax
    mass = np.array([np.abs(orc.transpoly_fit(row, 6)[1]).sum() for row in r[::40]]) / np.abs(r[::40]).max(axis=1)
    cond = 8 * np.finfo(np.float32).eps * mass.max()
    print(f"delta-Cp at BASELINE scale (F={F}, {int(valid.sum())} nodes): of operand scale: product vs float64 model "
          f"{e_exact.max():.2e}, product vs float-QR oracle {e_ref.max():.2e}, float-QR oracle vs float64 model {noise.max():.2e}; "
          f"of max|Cp| (SURVEY metric): product vs oracle max {ns_gpu.max():.2e} median {np.median(ns_gpu):.2e}, "
          f"oracle vs float64 model max {ns_orc.max():.2e} median {np.median(ns_orc):.2e}; nodes above 1e-5 of max|Cp|: "
          f"{int((ns_gpu > 1e-5).sum())} (oracle vs float64: {int((ns_orc > 1e-5).sum())})")
    assert e_exact.max() <= 1e-6 + cond
    assert np.all(e_ref <= 1e-5 + noise + cond)
    # rows of nodes the camera does not see / with a zero sample: NaN-ness must agree
    inval = ~valid & (cov[sel] != 0)
    assert np.array_equal(np.isnan(pt[sel][inval]), np.isnan(p_ref[inval]))


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_torchrun_two_processes(up, gpu, tmp_path, exchange):
    """exchange = nccl: the reference's structure (frame-major phase 1, local transpose into send blocks, grouped
    ncclSend / ncclRecv, reassembly; ncclAllReduce of the sums) through the same check."""
    if gpu < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1",
           "--frames", "1024", "--nodes", "60000", "--height", "512", "--width", "512", "--targets", "12",
           "--e2e-steps", "0", "--cpu-seconds", "0", "--check", "--exchange", exchange]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["n_gpus"] == 2 and out["parity_checked"] is True
    for c in out["parity"]["ranks"]:
        assert c["intensity_mismatches"] == 0 and c["avg_mismatches"] == 0 and c["rms_mismatches"] == 0
        assert c["intensity_values_checked"] > 1000
