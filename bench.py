#!/usr/bin/env python
"""bench.py -- cine -> surface delta-Cp throughput of the psp_process frame chain on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one whole psp_process job of BASELINE.json's configs[1] per GPU: 1 camera,
1024x1024 12-bit packed frames, 20 000 frames per GPU onto a 500 000-node grid
(weak scaling: F_total = 20 000 x N GPUs, nodes fixed): decode, hot-pixel fix, affine warp
(registration result supplied; the ECC solve is not part of either arm), fiducial patch,
projection, sum/sum-sq, frame-major -> node-major transpose (all-to-all when N > 1), detrend +
gain + delta-Cp.  One JSON line on stdout (rank 0).

  value   : frames/s with the packed frames already resident in HBM (device time, CUDA events
            on the library's stream, max over ranks)
  e2e     : frames/s through the C ABI with HOST buffers: pinned-host packed frames H2D every
            step, intensity_transpose + pressure_transpose D2H every step
  roofline: dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline: the CPU oracle (C port of the reference's algorithm, all host threads) on a
            bounded sample of the same workload (the single-GPU line only)
  parity  : --check (default on, outside the timed region): every rank compares sampled rows of what it holds after
            the last timed step with the CPU oracle (intensity_transpose, avg, rms, gain bit for bit; delta-Cp by the
            criterion of DESIGN.md section 4); "parity_checked" false and exit status != 0 on a mismatch
  --config {1,2,3} selects BASELINE.json's configs[1..3]; --exchange nccl the NCCL arm; --registration pixel the ECC line
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(height=1024, width=1024, frames_per_gpu=20000, nodes=500_000, cams=1, targets=32,
           distinct_frames=128, degree=6)

# --config N: BASELINE.json configs[N] as far as one box can hold it (the default, and the one `metric` is quoted on,
# is configs[1]; configs[0] is the CPU-runnable case of the test-suite, configs[4] is scripts/sweep_projection.py).
CONFIGS = {
    1: dict(),
    # 4 cameras with multi-camera blending (weighted projections, 40 % of the nodes unseen per camera), 1 M nodes;
    # 50 000 frames over 8 GPUs = 6 250 frames per GPU (per-GPU work fixed: the same slice on fewer GPUs)
    2: dict(cams=4, nodes=1_000_000, frames_per_gpu=6250,
            name="configs[2]: 4 cameras {H}x{W} 12-bit packed with blending, {F} frames/GPU onto {N}-node grid"),
    # multi-zone structured grid: 2 x the reference's 309 062-node test grid, 6 000 seam groups (overlap remap =
    # P3DModel::adjust_solution), 64 fiducial targets (65 patch clusters), degree-6 detrend
    3: dict(nodes=618_124, overlap_groups=6000, targets=64,
            name="configs[3]: multi-zone structured grid ({N} nodes, {G} seam groups), {K} patch clusters, detrend 6, "
                 "1 camera {H}x{W} 12-bit packed, {F} frames/GPU"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------
def build_workload(args, synth):
    H, W, N, C = args.height, args.width, args.nodes, args.cams
    t0 = time.time()
    frames = [synth.make_frames_fast(args.distinct, H, W, seed=1 + 10 * c) for c in range(C)]
    packed = [synth.pack_12bit(f.reshape(args.distinct, -1)) for f in frames]
    csr = [synth.make_projection(N, H, W, kind=args.csr, seed=1 + c, skipped_frac=0.02 if C == 1 else 0.4, weights=C > 1)
           for c in range(C)]
    patches = []
    nclusters = 0
    for c in range(C):
        bounds, internal = synth.make_patches(H, W, n_targets=args.targets, seed=3 + c)
        patches.append(synth.flatten_patches(bounds, internal))
        nclusters = len(bounds)
    remap = synth.overlap_src_index(N, synth.make_overlap(N, args.overlap_groups, seed=4)) if args.overlap_groups else None
    cal, qbar, ps, steady, temp = synth.tunnel_conditions(N)
    log(f"[bench] workload built in {time.time() - t0:.1f}s: {C} camera(s), {args.distinct} distinct frames {H}x{W}, "
        f"N={N}, nnz={[int(c[1].size) for c in csr]}, clusters={nclusters}, seam groups={args.overlap_groups}")
    return dict(frames=frames, packed=packed, csr=csr, patches=patches, remap=remap, cal=cal, qbar=qbar, ps=ps,
                steady=steady, temp=temp, nclusters=nclusters)


def oracle_patches(orc, wl, args):
    """The patch lists as the oracle takes them (one object per camera), or None."""
    if not args.targets:
        return None
    out = []
    for bo, bx, by, io, ix, iy in wl["patches"]:
        pobj = orc.Patches.__new__(orc.Patches)
        pobj.n, pobj.bounds_off, pobj.internal_off = bo.size - 1, bo, io
        pobj.bx, pobj.by, pobj.ix, pobj.iy = bx, by, ix, iy
        out.append(pobj)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return None
        hi = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa(local):
    """Pin this rank's host threads (and so, by first touch, its pinned buffers) to the NUMA node of its GPU: with
    every rank on node 0 the e2e arm of 8 GPUs moved 1.5x the bytes of one.  Returns a description."""
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                           capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = q[-12:] if len(q) >= 12 else q                     # 0000:17:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return {"numa_node": None, "note": "no NUMA affinity reported for the GPU"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus), "pci": bus}
    except (OSError, ValueError, subprocess.SubprocessError) as e:
        return {"numa_node": None, "note": f"not bound: {e}"}


def dist_setup(n_gpus):
    """torch.distributed is plumbing only: handle exchange + barriers."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        log(f"[bench] rank {rank}: host affinity {bind_to_gpu_numa(local)}")
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
        return rank, world, local, dist
    if n_gpus > 1:
        log("[bench] --gpus > 1 needs torchrun (one process per GPU); running rank 0 of 1")
    return 0, 1, 0, None


def barrier(dist):
    """Host-side rendezvous through the CPU (gloo) side of the process group: the data path has
    no NCCL collective (peer stores / peer reads inside the library's own kernels)."""
    if dist is not None:
        import torch
        t = torch.zeros(1, dtype=torch.int32)
        dist.all_reduce(t)


def allmax(dist, v):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def configure(up, wl, args, rank, world, local, capacity, dist):
    F_total = args.frames * world
    g = up.PspGpu(args.cams, args.nodes, F_total, device=local, rank=rank, n_ranks=world,
                  frame_capacity=capacity, batch_frames=args.batch)
    reg = {"given": up.REG_GIVEN, "none": up.REG_NONE, "pixel": up.REG_PIXEL}[args.registration]
    for c in range(args.cams):
        g.set_camera(c, args.width, args.height)
        g.set_projection(c, *wl["csr"][c])
    if wl["remap"] is not None:
        g.set_overlap_remap(wl["remap"])
    g.set_options(registration=reg, interp=up.INTERP_LINEAR,
                  patcher=up.PATCH_POLYNOMIAL if args.targets else up.PATCH_NONE)
    from upsp_b200 import synth
    for c in range(args.cams):
        if args.targets:
            g.set_patches(c, *wl["patches"][c])
        if reg == up.REG_GIVEN:
            g.set_warp_matrices(c, 0, synth.make_warps(g.n_frames, seed=5 + rank + 100 * c))
        if reg == up.REG_PIXEL:
            g.set_reference_frame(c, wl["frames"][c][0])
    if world > 1 and args.exchange == "nccl":
        # measured comparison: frame-major phase 1 + local transpose + grouped ncclSend / ncclRecv (no peer mappings)
        import torch
        g.set_exchange(up.XCHG_NCCL)
        uid = torch.frombuffer(bytearray(up.nccl_unique_id() if rank == 0 else bytes(up.NCCL_ID_BYTES)), dtype=torch.uint8).clone()
        dist.broadcast(uid, src=0)
        g.nccl_init(bytes(uid.numpy().tobytes()))
    elif world > 1:
        import torch
        h = torch.frombuffer(bytearray(g.ipc_export()), dtype=torch.uint8).clone()
        allh = [torch.empty_like(h) for _ in range(world)]
        dist.all_gather(allh, h)
        g.ipc_import(b"".join(bytes(x.numpy().tobytes()) for x in allh))
    return g


def run_step_resident(g, wl, args, dist, trace=None):
    t = [time.perf_counter()]
    g.reset_run()
    g.process_frames(0, g.n_frames)
    if dist is not None:
        g.sync()
        t.append(time.perf_counter())
        barrier(dist)            # MPI_Barrier before the reduce (psp_process.cpp:1859)
    t.append(time.perf_counter())
    g.finish_phase1()
    g.transpose()
    t.append(time.perf_counter())
    if dist is not None:
        barrier(dist)            # psp_process.cpp:2034-2035
    t.append(time.perf_counter())
    g.phase2(wl["cal"], wl["qbar"], wl["ps"], wl["steady"], wl["temp"], args.degree)
    t.append(time.perf_counter())
    if trace is not None:
        trace.append([round((b - a) * 1e3, 2) for a, b in zip(t[:-1], t[1:])])


# ------------------------------------------------------------------------------------------
def parity_check(g, wl, args, rank, world, dist, n_rows=192, frames_per_rank=24, n_rows_p2=48):
    """--check: after the last timed step every rank compares what IT holds with the CPU oracle
    (oracle/ = checker only), on the real multi-process path:
      1. intensity_transpose[sampled local nodes, sampled frames of EVERY rank's frame slice] bit-exact
         vs the oracle's phase 1 of those frames (decode, hot pixels, warp, patch, projection; the
         columns written by the peers over NVLink are checked on the rank that received them);
      2. avg / rms of the sampled nodes == finals of the sums of the rank's FULL rows, summed per
         rank slice in rank order like the device (a checksum over every column of the row: with
         unit projection values the sums are exact integers);
      3. pressure_transpose rows of a subset == the oracle's phase 2 on the same intensity rows:
         max_f |dCp| / max_f |Cp_ref| per node (the north-star metric) and the
         operand-scale figure of DESIGN.md section 4.
    Reference semantics: psp_process.cpp:707-771, 1866-1872.  Returns a dict; "ok" False fails the run."""
    from oracle import oracle as orc
    from upsp_b200 import synth
    orc.build()
    orc.set_num_threads(max(1, host_threads() // world))
    H, W, N, D = args.height, args.width, args.nodes, args.distinct
    F_local, F_total = args.frames, args.frames * world
    nl, n0 = g.n_local_nodes, g.first_node
    rng = np.random.default_rng(4242 + rank)
    rows = np.unique(np.concatenate([[0, nl - 1], rng.integers(0, nl, n_rows)])).astype(np.int64)
    itr = np.empty((rows.size, F_total), np.float32)
    ptr = np.empty((rows.size, F_total), np.float32)
    one = np.empty((1, F_total), np.float32)
    for i, r in enumerate(rows):
        g.read_intensity_transpose(int(r), 1, out=one)
        itr[i] = one[0]
        g.read_pressure_transpose(int(r), 1, out=one)
        ptr[i] = one[0]
    avg, rms, cov = g.read_phase1_stats()
    rms2, avg2, gain = g.read_phase2_stats()
    gl = rows + n0                                   # global node ids
    res = {"rank": rank, "rows": int(rows.size)}

    # ---- 1. sampled columns of every rank's slice through the oracle's phase 1 (all cameras; the overlap remap, when
    # there is one, is applied on the sampled rows: row n of the result is row src[n] of the projection)
    C = args.cams
    src = wl["remap"][gl] if wl["remap"] is not None else gl
    subs = []
    for c in range(C):
        rowptr, col, val = wl["csr"][c]
        sub_rowptr = np.zeros(rows.size + 1, np.int32)
        cnt = (rowptr[src + 1] - rowptr[src]).astype(np.int32)
        sub_rowptr[1:] = np.cumsum(cnt)
        idx = np.concatenate([np.arange(rowptr[n], rowptr[n + 1]) for n in src]) if cnt.sum() else np.zeros(0, np.int64)
        subs.append((sub_rowptr, col[idx].astype(np.int32), val[idx].astype(np.float32)))
    pobjs = oracle_patches(orc, wl, args)
    B = args.batch if args.batch > 0 else 256
    bad_cols, n_cols = 0, 0
    for r in range(world):
        # frames of rank r's slice: the first / last, batch edges, random ones
        loc = np.unique(np.concatenate([[0, 1, F_local - 1, min(B - 1, F_local - 1), min(B, F_local - 1)],
                                        rng.integers(0, F_local, frames_per_rank)])).astype(np.int64)
        warps = ([synth.make_warps(F_local, seed=5 + r + 100 * c) for c in range(C)] if args.registration == "given" else None)
        fr = [orc.unpack_12bit_frames(wl["packed"][c][loc % D]).reshape(loc.size, H, W) for c in range(C)]
        for first, sel in ((0, (loc == 0) & (r == 0)), (1, ~((loc == 0) & (r == 0)))):
            if not sel.any():
                continue
            it, _, _ = orc.phase1([f[sel] for f in fr], subs, first_frame=first,
                                  warp=[w[loc[sel]] for w in warps] if warps is not None else None, interp=1,
                                  patches=pobjs)
            got = itr[:, r * F_local + loc[sel]].T           # [frames, rows]
            same = (it.view(np.uint32) == np.ascontiguousarray(got).view(np.uint32)) | (np.isnan(it) & np.isnan(got))
            bad_cols += int((~same).sum())
            n_cols += int(same.size)
    res["intensity_values_checked"] = n_cols
    res["intensity_mismatches"] = bad_cols

    # ---- 2. sums of the full rows, per rank slice in rank order (the device's order)
    s = np.zeros(rows.size)
    q = np.zeros(rows.size)
    for r in range(world):
        blk = itr[:, r * F_local:(r + 1) * F_local]
        s += np.cumsum(blk.astype(np.float64), axis=1)[:, -1]
        q += np.cumsum((blk * blk).astype(np.float64), axis=1)[:, -1]
    a_ref, r_ref = orc.phase1_finals(s, q, F_total)
    a_got, r_got = avg[gl], rms[gl]
    eq = lambda x, y: (x.view(np.uint32) == y.view(np.uint32)) | (np.isnan(x) & np.isnan(y))
    res["avg_mismatches"] = int((~eq(a_ref, a_got)).sum())
    res["rms_mismatches"] = int((~eq(r_ref, r_got)).sum())

    # ---- 3. phase 2 of a subset of the rows through the oracle: in the reference's float arithmetic (restated
    # float QR) and with the float64 least-squares fit.  Criterion of DESIGN.md section 4 / tests/test_gpu_parity.py:
    # product vs float64 model <= 1e-6 (+ the precision to which the float design matrix defines the fit), product vs
    # float-QR oracle <= 1e-5 + that oracle's own distance from the float64 model, all relative to the operands of
    # r - fit; the north-star figure max_f|dCp| / max_f|Cp_ref| is reported for both.
    k = np.linspace(0, rows.size - 1, min(n_rows_p2, rows.size)).astype(np.int64)
    oargs = (itr[k], a_got[k], cov[gl[k]], wl["steady"][gl[k]], wl["temp"][gl[k]], wl["cal"], wl["qbar"], wl["ps"], args.degree)
    p_ref, rms2_ref, avg2_ref, gain_ref = orc.phase2(*oargs)
    p_exact = orc.phase2(*oargs, exact_fit=True)[0]
    valid = (cov[gl[k]] != 0) & np.all(np.isfinite(itr[k]) & (itr[k] != 0), axis=1)
    res["cp_rows"] = int(valid.sum())
    res["gain_mismatches"] = int((~eq(gain_ref[valid], gain[rows[k]][valid])).sum())
    cp_ok = True
    if valid.any():
        Kn = np.abs(gain_ref[valid]).astype(np.float64) * 144.0 / float(wl["qbar"])
        r = (a_got[k][valid, None] / itr[k][valid]).astype(np.float32)
        scale = Kn * np.abs(r).max(axis=1)
        err = lambda a, b: np.abs(a[valid] - b[valid]).max(axis=1)
        e_exact, e_ref, noise = err(ptr[k], p_exact) / scale, err(ptr[k], p_ref) / scale, err(p_ref, p_exact) / scale
        cpmax = np.abs(p_ref[valid]).max(axis=1)
        mass = np.array([np.abs(orc.transpoly_fit(row, args.degree)[1]).sum() for row in r]) / np.abs(r).max(axis=1)
        cond = 8 * np.finfo(np.float32).eps * mass
        res["cp_err_vs_float64_model_of_operands"] = float(e_exact.max())
        res["cp_err_vs_floatqr_oracle_of_operands"] = float(e_ref.max())
        res["floatqr_oracle_vs_float64_model_of_operands"] = float(noise.max())
        res["cp_err_vs_oracle_of_max_cp"] = float((err(ptr[k], p_ref) / cpmax).max())          # north-star metric
        res["oracle_vs_float64_model_of_max_cp"] = float((err(p_ref, p_exact) / cpmax).max())
        cp_ok = bool(np.all(e_exact <= 1e-6 + cond) and np.all(e_ref <= 1e-5 + noise + cond))
    inval = ~valid & (cov[gl[k]] != 0)
    nanrows_same = bool(np.array_equal(np.isnan(ptr[k][inval]).all(axis=1), np.isnan(p_ref[inval]).all(axis=1)))
    res["ok"] = (bad_cols == 0 and res["avg_mismatches"] == 0 and res["rms_mismatches"] == 0 and
                 res["gain_mismatches"] == 0 and nanrows_same and cp_ok)
    return res


def bench_b200(args):
    import upsp_b200 as up
    from upsp_b200 import synth
    up.build.build()
    rank, world, local, dist = dist_setup(args.gpus)
    if up.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible (there is no CPU fallback)")
    wl = build_workload(args, synth)
    P = args.height * args.width
    N = args.nodes
    F_local = args.frames
    F_total = F_local * world

    # ---------------- device-resident arm
    g = configure(up, wl, args, rank, world, local, 0, dist)
    D = args.distinct
    for o in range(0, g.n_frames, D):
        n = min(D, g.n_frames - o)
        for c in range(args.cams):
            g.push_frames(c, wl["packed"][c][:n], up.PIX_PACKED12, o, n)
    g.sync()
    for _ in range(args.warmup):
        run_step_resident(g, wl, args, dist)
    g.sync()
    # rank 0 forks nvidia-smi BEFORE the rendezvous: the fork of this process (tens of GB of pinned host memory mapped)
    # takes ~20 ms, and a rank that starts its timer that much earlier only waits for rank 0 at the first in-step barrier,
    # which the max-over-ranks device time then charges to the steps (7 ms per step at --steps 3)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier(dist)
    # per-kernel CUDA events in the LAST timed step only: every 8th batch runs un-overlapped (the
    # library serialises a sampled batch so that event durations are the kernels' own) + phase 2
    g.set_kernel_sampling(0)
    l0 = g.launch_count()
    stage = np.zeros(4)
    g.timer_start()
    t_wall = time.time()
    trace = []
    for _ in range(args.steps):
        if _ == args.steps - 1:
            g.set_kernel_sampling(8)
        run_step_resident(g, wl, args, dist, trace)
        stage += [g.stage_ms(i) for i in range(4)]
        if _ == args.steps - 1:
            kms = [g.kernel_ms(k) for k in range(7)]     # of the last timed step
    ms_dev = g.timer_stop()
    pmode = g.projection_mode()          # 0: k_project_fused4 (global taps), 1: TMA boxes of decoded frames, 2: of packed frames
    row_b = g.row_bytes()                # 2: node-major rows stored as 16-bit integers (unit projection values), else 4
    g.sync()
    barrier(dist)
    wall_ms = (time.time() - t_wall) * 1e3
    clocks = sampler.stop() if sampler else None
    log(f"[bench] rank {rank} host timeline per step (ms; process[, barrier], finish+transpose[, barrier], phase2): {trace}")
    launches = g.launch_count() - l0
    ms_dev = allmax(dist, ms_dev)
    stage /= args.steps
    stage = np.array([allmax(dist, float(s)) for s in stage])
    value = F_total * args.steps / (ms_dev * 1e-3)
    check = None
    if args.check:
        check = parity_check(g, wl, args, rank, world, dist)
        log(f"[bench] rank {rank} parity check: {check}")
        if dist is not None:
            allc = [None] * world
            dist.all_gather_object(allc, check)
        else:
            allc = [check]
        check = {"ok": all(c["ok"] for c in allc), "ranks": allc}
    g.close()
    del g

    # ---------------- end-to-end arm (host buffers through the C ABI)
    e2e = None
    if args.e2e_steps > 0 and args.exchange == "peer":      # the streamed column reads of the e2e arm need the fused exchange
        e2e = bench_e2e(up, wl, args, rank, world, local, dist)

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # per-kernel launches of one step, CUDA-event mean duration (sampled inside the timed region),
    # algorithmic bytes per launch
    B = args.batch if args.batch > 0 else 256     # library default (upsp_gpu_config.batch_frames = 0)
    nbatch = -(-F_local // B)
    # front end / projection kernel of the mode that ran, with their algorithmic bytes per launch:
    # mode 2 never writes decoded frames: the scan reads the packed bytes once, the projection reads them again
    knames = ["k_hot_scan12" if pmode == 2 else "k_unpack12_scan_p", "k_frame_prep", "k_warp_affine8_u16", "k_patch",
              "k_project_tma" if pmode else "k_project_fused4", "k_transpose_a2a", "k_phase2_sym"]
    kalg = [B * 1.5 * P if pmode == 2 else B * 3.5 * P, 0.0, B * 4.0 * P, 0.0,
            B * ((1.5 if pmode == 2 else 2.0) * P * args.cams + row_b * N), 8.0 * N * F_local,
            (row_b + 4.0) * (N / world) * F_total]
    klaunch = [nbatch * args.cams, nbatch, nbatch, nbatch * args.cams, nbatch, 1, 1]
    kernels = {}
    for nm, (ms, ns), ab, nl in zip(knames, kms, kalg, klaunch):
        if ns:
            kernels[nm] = {"mean_ms": round(ms, 4), "launches_per_step": nl, "ms_per_step": round(ms * nl, 3),
                           "alg_bytes_per_launch": ab, "gbs": round(ab / (ms * 1e-3) / 1e9, 1) if ab else None,
                           "sampled": ns}
    dom = max((k for k in kernels if kernels[k]["alg_bytes_per_launch"]), key=lambda k: kernels[k]["ms_per_step"])
    achieved = kernels[dom]["gbs"]
    # DRAM traffic of the dominant kernel per launch from the committed `ncu --set full` capture
    # (profiles/r01_traffic.json, written by scripts/ncu_summary.py at the same batch size), else null
    traffic = None
    for tf in ("r02_traffic.json", "r01_traffic.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", tf)))
            if tr.get(dom, {}).get("batch_frames") == B and tr[dom].get("nodes") == N:
                traffic = tr[dom]["dram_bytes_per_launch"]
                break
        except (OSError, ValueError):
            pass
    names = ["process_frames", "finish_phase1", "transpose", "phase2"]
    chain_bytes = (1.5 * P * args.cams + 20.0 * N) * F_local
    chain_gbs = chain_bytes / (ms_dev / args.steps * 1e-3) / 1e9
    out = {
        "metric": "frames/sec cine->surface Cp", "value": round(value, 1), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (u16 pixels, f64 accumulators)", "data": "synthetic",
        "config": bench_config(args, world),
        "intensity_row_bytes": row_b,
        "projection_mode": {0: "global taps (k_project_fused4)", 1: "TMA boxes of decoded u16 frames",
                            2: "TMA boxes of the packed 12-bit frames"}.get(pmode, str(pmode)),
        "stage_ms": {n: round(float(s), 3) for n, s in zip(names, stage)},
        "chain": {"algorithmic_bytes_per_frame": 1.5 * P * args.cams + 20.0 * N, "achieved_gbs": round(chain_gbs, 1),
                  "frac_of_peak": round(chain_gbs / peak, 4),
                  "note": "SURVEY 8d formula (frame read + 4N row write + 8N transpose + 8N phase 2); the fused "
                          "projection writes node-major rows directly, so the implementation moves 8N less"},
        "kernels": kernels,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak,
                     "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": kernels[dom]["alg_bytes_per_launch"],
                     "launch_ms": kernels[dom]["mean_ms"], "share_of_step": round(kernels[dom]["ms_per_step"] / (ms_dev / args.steps), 3)},
        "gpu_launches": int(launches), "wall_ms_per_step": round(wall_ms / args.steps, 3),
        "clocks": clocks, "e2e": e2e,
    }
    if check is not None:
        out["parity_checked"] = bool(check["ok"])
        out["parity"] = check
    if args.cpu_seconds > 0 and world == 1:       # a reported baseline of the single-GPU line only
        out["cpu_baseline"] = cpu_baseline(args, wl, budget_s=args.cpu_seconds)
    print(json.dumps(out), flush=True)
    if check is not None and not check["ok"]:
        raise SystemExit("bench.py --check: parity check FAILED (see the parity object of the JSON line)")


def bench_e2e(up, wl, args, rank, world, local, dist):
    """Same job, host buffers: every step pushes all packed frames from pinned host memory
    (H2D overlapped with processing through the library's copy stream) and reads
    intensity_transpose and pressure_transpose back to pinned host memory."""
    import torch
    N, F_local = args.nodes, args.frames
    F_total = F_local * world
    chunk = args.distinct
    g = configure(up, wl, args, rank, world, local, 2 * chunk, dist)
    fb = wl["packed"][0].shape[1]
    pin_in = [torch.empty((chunk, fb), dtype=torch.uint8, pin_memory=True) for _ in range(args.cams)]
    for c in range(args.cams):
        pin_in[c].numpy()[:] = wl["packed"][c][:chunk]
    rows = max(1, min(g.n_local_nodes, (256 << 20) // (F_total * 4)))     # 256 MB D2H staging
    pin_out = torch.empty((rows, F_total), dtype=torch.float32, pin_memory=True)
    # streamed output: intensity_transpose leaves in column blocks of `cb` frames while later
    # frames are still being pushed / processed (H2D and D2H share the full-duplex PCIe link)
    cb = 8 * chunk
    pin_cols = [torch.empty((g.n_local_nodes, cb), dtype=torch.float32, pin_memory=True) for _ in range(2)]

    def step():
        g.reset_run()
        nblk = 0
        for o in range(0, g.n_frames, chunk):
            n = min(chunk, g.n_frames - o)
            for c in range(args.cams):
                g.push_frames(c, pin_in[c].data_ptr(), up.PIX_PACKED12, o, n)
            g.process_frames(o, n)
            done = o + n
            if done % cb == 0 or done == g.n_frames:
                b0 = (done - 1) // cb * cb
                g.read_intensity_transpose_block_async(0, g.n_local_nodes, g.first_frame + b0, done - b0,
                                                       pin_cols[nblk % 2].data_ptr(), cb)
                nblk += 1
        if dist is not None:
            g.sync()
            barrier(dist)
        g.finish_phase1()
        g.transpose()
        if dist is not None:
            barrier(dist)
            # columns written by the peers (their frame slices): read after the barrier
            for r in range(world):
                if r == rank:
                    continue
                for lo in range(r * F_local, (r + 1) * F_local, cb):
                    n = min(cb, (r + 1) * F_local - lo)
                    g.read_intensity_transpose_block_async(0, g.n_local_nodes, lo, n, pin_cols[0].data_ptr(), cb)
        g.phase2(wl["cal"], wl["qbar"], wl["ps"], wl["steady"], wl["temp"], args.degree)
        g.wait_reads()
        for o in range(0, g.n_local_nodes, rows):
            g.read_raw("pressure_transpose", o, min(rows, g.n_local_nodes - o), pin_out.data_ptr())

    step()                                   # warm-up
    g.sync()
    barrier(dist)
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step()
    g.sync()
    barrier(dist)
    dt = allmax(dist, time.perf_counter() - t0)
    g.close()
    return {"value": round(F_total * args.e2e_steps / dt, 1), "unit": "frames/s",
            "h2d_bytes_per_step": int(fb) * F_local * world * args.cams,
            "d2h_bytes_per_step": 2 * 4 * N * F_total, "steps": args.e2e_steps,
            "ms_per_step": round(dt / args.e2e_steps * 1e3, 1),
            "note": "H2D of packed 12-bit frames from pinned host memory + D2H of intensity_transpose "
                    "(streamed in column blocks while later frames are processed) and pressure_transpose "
                    "(the two flat files the reference writes) inside the timed region"}


# ------------------------------------------------------------------------------------------
def host_threads():
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1: ignored on purpose)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_job(orc, wl, args, n_frames, seed_rank=0):
    """The reference's algorithm (CPU oracle port, OpenMP over frames / nodes like the
    reference) on the first n_frames frames of the same workload, starting from the PACKED 12-bit
    frames like the GPU arm.  Returns seconds per stage."""
    from upsp_b200 import synth
    C = args.cams
    idx = np.arange(n_frames) % wl["packed"][0].shape[0]
    pk = [wl["packed"][c][idx] for c in range(C)]
    pobjs = oracle_patches(orc, wl, args)
    warp = ([synth.make_warps(n_frames, seed=5 + seed_rank + 100 * c) for c in range(C)] if args.registration == "given" else None)
    t0 = time.perf_counter()
    fr = [orc.unpack_12bit_frames(p).reshape(n_frames, args.height, args.width) for p in pk]
    inten, s, q = orc.phase1(fr, wl["csr"], warp=warp, interp=1, patches=pobjs, remap=wl["remap"])
    avg, rms = orc.phase1_finals(s, q, n_frames, wl["remap"])
    cov = orc.coverage(wl["csr"], wl["remap"])
    t1 = time.perf_counter()
    itr = orc.global_transpose([inten], args.nodes, n_frames)[0]
    t2 = time.perf_counter()
    orc.phase2(itr, avg, cov, wl["steady"], wl["temp"], wl["cal"], wl["qbar"], wl["ps"], args.degree)
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2


def workload_name(args):
    """config.workload: the same string in both arms (the driver compares them)."""
    name = CONFIGS.get(args.config, {}).get("name")
    if name:
        return name.format(H=args.height, W=args.width, F=args.frames, N=args.nodes, G=args.overlap_groups,
                           K=args.targets + 1 if args.targets else 0)
    return (f"configs[1]: 1 camera {args.height}x{args.width} 12-bit packed, "
            f"{args.frames} frames/GPU onto {args.nodes}-node grid")


def cpu_baseline(args, wl, budget_s=20.0):
    from oracle import oracle as orc
    orc.build()
    threads = orc.set_num_threads(host_threads())
    # probe with a few frames, then size the sample to ~budget_s of CPU work
    p = cpu_job(orc, wl, args, 2 * threads)
    per_frame = sum(p) / (2 * threads)
    n = int(max(4 * threads, min(2048, budget_s / max(per_frame, 1e-6))))
    a, b, c = cpu_job(orc, wl, args, n)
    return {"value": round(n / (a + b + c), 2), "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{n} packed frames of the same workload (same frame size, grid, patches, warp); "
                      f"process-frames {a:.2f}s, transpose {b:.2f}s, phase2 {c:.2f}s",
            "note": "C port of the reference's algorithm (oracle/upsp_oracle.c), OpenMP over frames and "
                    "nodes as the reference; the reference's own C++ cannot be compiled here (DESIGN.md)"}


def bench_reference(args):
    """--impl reference: the reference's CPU algorithm on the host cores (oracle port; the
    reference itself needs OpenCV C++/Eigen/MPI/HDF5, none of which exist in this image).  Rank 0 only;
    all the host's cores whatever OMP_NUM_THREADS says (torchrun sets it to 1); each step is a bounded
    sample of the GPU arm's job: the first `--ref-frames` packed frames (default 2048), per-frame cost
    being flat in the frame count."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import upsp_b200  # noqa: F401  (synthetic workload generator only)
    from upsp_b200 import synth
    from oracle import oracle as orc
    orc.build()
    wl = build_workload(args, synth)
    threads = orc.set_num_threads(host_threads())
    n = args.ref_frames if args.ref_frames > 0 else min(2048, args.frames)
    for _ in range(args.warmup):
        cpu_job(orc, wl, args, max(threads, n // 8))
    t0 = time.perf_counter()
    parts = np.zeros(3)
    for _ in range(args.steps):
        parts += cpu_job(orc, wl, args, n)
    dt = time.perf_counter() - t0
    v = round(n * args.steps / dt, 2)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample = (f"each step = the first {n} packed frames of the job ({args.height}x{args.width}, N={args.nodes}); "
              f"process-frames {parts[0] / args.steps:.2f}s transpose {parts[1] / args.steps:.2f}s "
              f"phase2 {parts[2] / args.steps:.2f}s per step")
    print(json.dumps({
        "impl": "reference", "metric": "frames/sec cine->surface Cp", "value": v, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (u16 pixels, f64 accumulators)", "data": "synthetic",
        "config": bench_config(args, world),
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def bench_config(args, world):
    """The `config` object of the JSON line: identical in both arms."""
    return {"workload": workload_name(args), "frames_total": args.frames * world, "nodes": args.nodes,
            "cameras": args.cams, "registration": args.registration,
            "patch_clusters": args.targets + 1 if args.targets else 0, "seam_groups": args.overlap_groups,
            "csr": args.csr, "detrend_degree": args.degree, "batch_frames": args.batch, "exchange": args.exchange,
            "l2": "inputs (tens of GB of packed frames and of intensity rows per GPU) far exceed the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS),
                    help="BASELINE.json configs[N] (1 = the configuration the metric is quoted on)")
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU")
    ap.add_argument("--nodes", type=int, default=None)
    ap.add_argument("--cams", type=int, default=None)
    ap.add_argument("--overlap-groups", type=int, default=None, help="seam groups of a multi-zone grid (overlap remap)")
    ap.add_argument("--height", type=int, default=CFG["height"])
    ap.add_argument("--width", type=int, default=CFG["width"])
    ap.add_argument("--targets", type=int, default=None)
    ap.add_argument("--distinct", type=int, default=CFG["distinct_frames"])
    ap.add_argument("--degree", type=int, default=CFG["degree"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--csr", default="surface", choices=["surface", "random"])
    ap.add_argument("--registration", default="given", choices=["given", "none", "pixel"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="peer: exchange fused into the projection kernel (default); nccl: the reference's structure on NCCL")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget (0 = skip)")
    ap.add_argument("--ref-frames", type=int, default=0)
    ap.add_argument("--check", action=argparse.BooleanOptionalAction, default=True,
                    help="after the last timed step (outside the timed region) compare sampled outputs of every rank "
                         "with the CPU oracle; rc != 0 on mismatch, \"parity_checked\" in the JSON line (--no-check: skip)")
    args = ap.parse_args()
    preset = dict(CFG, overlap_groups=0)
    preset.update({k: v for k, v in CONFIGS[args.config].items() if k != "name"})
    for key, attr in (("frames_per_gpu", "frames"), ("nodes", "nodes"), ("cams", "cams"), ("overlap_groups", "overlap_groups"),
                      ("targets", "targets")):
        if getattr(args, attr) is None:
            setattr(args, attr, preset[key])
    if args.warmup < 3 and args.impl == "b200":
        log("[bench] note: fewer than 3 warm-up steps requested")
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_b200(args)


if __name__ == "__main__":
    main()
