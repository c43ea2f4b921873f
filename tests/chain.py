"""Shared drivers for the parity tests: run the same synthetic psp_process case through the
CPU oracle and through the CUDA library (C ABI), rank by rank."""
from __future__ import annotations

import numpy as np


class Case:
    """A complete synthetic psp_process input (SURVEY 8d config-1 style)."""

    def __init__(self, synth, *, n_cams=1, n_nodes=3000, n_frames=48, height=96, width=128,
                 registration=False, interp=1, patches=False, overlap=False, kind="surface",
                 weights=False, multi_nnz=0, seed=0, degree=6, fmt="u16", overlap_pair=False, jitter=True, texture=250.0, filter_kind=0, filter_size=0):
        self.C, self.N, self.F, self.H, self.W = n_cams, n_nodes, n_frames, height, width
        self.interp, self.degree, self.fmt = interp, degree, fmt
        self.filter_kind, self.filter_size = filter_kind, filter_size
        self.frames = [synth.make_frames(n_frames, height, width, seed=seed + 10 * c, jitter=jitter, texture=texture)[0]
                       for c in range(n_cams)]
        if multi_nnz:
            self.csr = [synth.make_multi_nnz_projection(n_nodes, height, width, multi_nnz, seed=seed + 20 + c)
                        for c in range(n_cams)]
        else:
            self.csr = [synth.make_projection(n_nodes, height, width, kind=kind, seed=seed + 20 + c,
                                              skipped_frac=0.02 if n_cams == 1 else 0.4,
                                              weights=weights or n_cams > 1)
                        for c in range(n_cams)]
        self.warp = ([synth.make_warps(n_frames, seed=seed + 30 + c) for c in range(n_cams)]
                     if registration else None)
        self.patch_lists = ([synth.make_patches(height, width, n_targets=6, seed=seed + 40 + c,
                                                overlap_pair=overlap_pair) for c in range(n_cams)]
                            if patches else None)
        self.overlap = synth.make_overlap(n_nodes, 40, seed=seed + 50) if overlap else None
        self.cal, self.qbar, self.ps, self.steady, self.temp = synth.tunnel_conditions(n_nodes, seed + 60)
        self.synth = synth


def run_oracle(orc, case: Case, n_ranks=1, exact_fit=False):
    """Reference flow: phase-1 loop per rank, reduce, finals, global transpose, phase 2."""
    N, F = case.N, case.F
    remap = orc.overlap_remap(N, case.overlap) if case.overlap is not None else None
    patches = ([orc.Patches(b, i) for (b, i) in case.patch_lists] if case.patch_lists else None)
    fs, fe = orc.apportion(F, n_ranks)
    ns, ne = orc.apportion(N, n_ranks)
    sum_ = np.zeros(N)
    sumsq = np.zeros(N)
    inten = []
    for r in range(n_ranks):
        sl = slice(int(fs[r]), int(fs[r] + fe[r]))
        it, s, q = orc.phase1([f[sl] for f in case.frames], case.csr, first_frame=int(fs[r]),
                              warp=[w[sl] for w in case.warp] if case.warp else None,
                              interp=case.interp, patches=patches, remap=remap,
                              filter_kind=case.filter_kind, filter_size=case.filter_size)
        inten.append(it)
        sum_ += s
        sumsq += q
    avg, rms = orc.phase1_finals(sum_, sumsq, F, remap)
    cov = orc.coverage(case.csr, remap)
    itrans = orc.global_transpose(inten, N, F)
    out = dict(intensity=np.concatenate(inten, 0), avg=avg, rms=rms, coverage=cov,
               itrans=np.concatenate(itrans, 0))
    pt, r2, a2, g2 = [], [], [], []
    for r in range(n_ranks):
        sl = slice(int(ns[r]), int(ns[r] + ne[r]))
        p, rr, aa, gg = orc.phase2(itrans[r], avg[sl], cov[sl], case.steady[sl], case.temp[sl],
                                   case.cal, case.qbar, case.ps, case.degree, exact_fit=exact_fit)
        pt.append(p)
        r2.append(rr)
        a2.append(aa)
        g2.append(gg)
    out.update(ptrans=np.concatenate(pt, 0), rms2=np.concatenate(r2), avg2=np.concatenate(a2),
               gain=np.concatenate(g2))
    return out


def setup_ctx(up, orc_mod, case: Case, rank=0, n_ranks=1, device=0, batch_frames=0,
              frame_capacity=0, alias=True, keep_frame_major=False):
    """Create + configure one rank's context through the C ABI (no oracle involvement:
    orc_mod is only used for the pure-index helpers pack_12bit / overlap_remap)."""
    g = up.PspGpu(case.C, case.N, case.F, device=device, rank=rank, n_ranks=n_ranks,
                  batch_frames=batch_frames, frame_capacity=frame_capacity,
                  pressure_aliases_intensity=alias, keep_frame_major=keep_frame_major)
    for c in range(case.C):
        g.set_camera(c, case.W, case.H)
        g.set_projection(c, *case.csr[c])
    if case.overlap is not None:
        g.set_overlap_remap(orc_mod.overlap_remap(case.N, case.overlap))
    g.set_options(registration=up.REG_GIVEN if case.warp else up.REG_NONE, interp=case.interp,
                  patcher=up.PATCH_POLYNOMIAL if case.patch_lists else up.PATCH_NONE)
    if case.filter_kind:
        g.set_filter(case.filter_kind, case.filter_size)
    if case.patch_lists:
        for c in range(case.C):
            g.set_patches(c, *case.synth.flatten_patches(*case.patch_lists[c]))
    sl = slice(g.first_frame, g.first_frame + g.n_frames)
    if case.warp:
        for c in range(case.C):
            g.set_warp_matrices(c, 0, case.warp[c][sl])
    return g, sl


def push_all(up, orc_mod, g, case: Case, sl, chunk=None):
    nloc = sl.stop - sl.start
    chunk = chunk or max(nloc, 1)
    for o in range(0, nloc, chunk):
        n = min(chunk, nloc - o)
        for c in range(case.C):
            fr = case.frames[c][sl.start + o: sl.start + o + n]
            if case.fmt == "p12":
                packed = np.stack([orc_mod.pack_12bit(f) for f in fr]) if n else np.zeros((0, 1), np.uint8)
                g.push_frames(c, packed, up.PIX_PACKED12, o, n)
            else:
                g.push_frames(c, fr, up.PIX_U16, o, n)
        g.process_frames(o, n)


def run_gpu(up, orc_mod, case: Case, **kw):
    """Single-rank run through the C ABI.  Returns the same dict as run_oracle."""
    g, sl = setup_ctx(up, orc_mod, case, **kw)
    cap = kw.get("frame_capacity", 0)
    push_all(up, orc_mod, g, case, sl, chunk=cap if cap else None)
    g.finish_phase1()
    out = dict(intensity=g.read_intensity() if kw.get("keep_frame_major") else None)
    out["avg"], out["rms"], out["coverage"] = g.read_phase1_stats()
    g.transpose()
    out["itrans"] = g.read_intensity_transpose()
    g.phase2(case.cal, case.qbar, case.ps, case.steady, case.temp, case.degree)
    out["ptrans"] = g.read_pressure_transpose()
    out["rms2"], out["avg2"], out["gain"] = g.read_phase2_stats()
    out["launches"] = g.launch_count()
    g.close()
    return out


def same_bits(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    nan_a = (a & 0x7fffffff) > 0x7f800000
    nan_b = (b & 0x7fffffff) > 0x7f800000
    return bool(np.all((a == b) | (nan_a & nan_b)))


def degenerate_nodes(ref):
    """Covered nodes whose intensity history holds a zero / non-finite sample (e.g. a node whose
    pixel is warped in from outside the frame: BORDER_CONSTANT 0).  avg / 0 = inf poisons the
    reference's whole row (NaN after the QR solve); there is nothing to compare but NaN-ness."""
    it = ref["itrans"]
    return (ref["coverage"] != 0) & ~np.all(np.isfinite(it) & (it != 0), axis=1)


def cp_errors(case: Case, ref, got):
    """Per-node error of the delta-Cp time histories.
    Returns (err_operand, err_cp):
      err_operand[n] = max_f |dCp| / (K_n * max_f |r|)  -- relative to the operands of the
                       cancelling subtraction (r - fit), K_n = |gain|*144/qbar;  this is the
                       1e-5 (fp32) criterion of BASELINE.json, see DESIGN.md "tolerances".
      err_cp[n]      = max_f |dCp| / max_f |Cp_ref|      -- relative to the signal itself."""
    valid = (ref["coverage"] != 0) & ~degenerate_nodes(ref)
    d = np.abs(got["ptrans"][valid] - ref["ptrans"][valid]).max(axis=1)
    K = np.abs(ref["gain"][valid]).astype(np.float64) * 144.0 / float(case.qbar)
    r = ref["avg"][valid, None] / ref["itrans"][valid]
    err_operand = d / (K * np.abs(r).max(axis=1))
    err_cp = d / np.abs(ref["ptrans"][valid]).max(axis=1)
    return err_operand, err_cp


def monomial_mass(orc, ref):
    """sum_c |coef_c| of the reference's degree-6 monomial fit per valid node, divided by
    max|r|.  The reference's design matrix A[f,c] = (float)pow(f/F, c) is only defined to float
    precision, so fitted values are only defined to ~eps32 * sum_c|coef_c| (each column entry
    carries a relative rounding error of up to 6e-8 that is multiplied by its coefficient);
    rows with outliers have large alternating coefficients."""
    valid = (ref["coverage"] != 0) & ~degenerate_nodes(ref)
    r = (ref["avg"][valid, None] / ref["itrans"][valid]).astype(np.float32)
    mass = np.array([np.abs(orc.transpoly_fit(row, 6)[1]).sum() for row in r])
    return mass / np.abs(r).max(axis=1)
