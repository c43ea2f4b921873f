"""BASELINE.json configs[4] / SURVEY 8d config 5: the projection (K5) alone, swept over frame size, frame batch and
CSR shape, against the HBM roofline.

    python scripts/sweep_projection.py [--out profiles/r02_projection_sweep.csv]

Per point one psp_process context through the C ABI (u16 frames resident in HBM, registration none, patcher none) whose
projection kernel is bracketed by CUDA events on every batch (upsp_gpu_set_kernel_sampling(1)):
  * nnz/row 1, fused  : the fused projection (node-major rows written directly: projection + transpose);
  * nnz/row 1, plain  : k_project_ell1 (frame-major rows, the reference's project_frame layout, projection.ipp:884-908);
  * nnz/row 4, 9      : k_project_csr (general CSR, frame-major rows);
  * csr "surface" (mesh-like locality) and "random" (uniformly random columns: worst-case gather).
Algorithmic bytes per launch = batch * (2 P + 4 N) (u16 frame read + f32 row write; the CSR arrays, 8 B per entry, are
read once per launch and added for nnz > 1).  Fraction = achieved GB/s over MEASURED_PEAKS.json's hbm_gbs."""
import argparse
import csv
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_projection_sweep.csv"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import upsp_b200 as up
    from upsp_b200 import synth
    up.build.build()
    peak = 6553.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    sizes = [(512, 512), (1024, 1024), (1920, 2048)] if not args.quick else [(512, 512)]
    batches = [1, 8, 32, 128, 256] if not args.quick else [8, 128]
    rows = []
    for (H, W) in sizes:
        P = H * W
        N = P // 2                                   # one node per two pixels, as in configs[1]
        frames = synth.make_frames_fast(32, H, W, seed=1)
        for kind in ("surface", "random"):
            for nnz in (1, 4, 9):
                if nnz == 1:
                    csr = synth.make_projection(N, H, W, kind=kind, seed=1)
                elif kind == "surface":       # a k x k stencil of weighted pixels around every node's pixel, nodes in raster order
                    rng = np.random.default_rng(2)
                    k = 2 if nnz == 4 else 3
                    base = np.sort(rng.integers(0, P, N))
                    by, bx = base // W, base % W
                    dy, dx = np.divmod(np.arange(nnz), k)
                    yy = np.clip(by[:, None] + dy[None, :] - (k - 1) // 2, 0, H - 1)
                    xx = np.clip(bx[:, None] + dx[None, :] - (k - 1) // 2, 0, W - 1)
                    rowptr = (np.arange(N + 1, dtype=np.int64) * nnz).astype(np.int32)
                    csr = (rowptr, (yy * W + xx).astype(np.int32).ravel(), rng.uniform(0.05, 0.6, N * nnz).astype(np.float32))
                else:
                    rng = np.random.default_rng(3)
                    rowptr = (np.arange(N + 1, dtype=np.int64) * nnz).astype(np.int32)
                    csr = (rowptr, rng.integers(0, P, N * nnz).astype(np.int32), rng.uniform(0.1, 1.0, N * nnz).astype(np.float32))
                for fused in ((True, False) if nnz == 1 else (False,)):
                    for B in batches:
                        F = max(4 * B, 64)
                        g = up.PspGpu(1, N, F, batch_frames=B, keep_frame_major=not fused)
                        g.set_camera(0, W, H)
                        g.set_projection(0, *csr)
                        g.set_options(registration=up.REG_NONE, patcher=up.PATCH_NONE)
                        for o in range(0, F, 32):
                            n = min(32, F - o)
                            g.push_frames(0, frames[:n], up.PIX_U16, o, n)
                        g.process_frames(0, F)          # warm-up
                        g.sync()
                        g.reset_run()
                        g.set_kernel_sampling(1)
                        g.process_frames(0, F)
                        g.sync()
                        ms, ns = g.kernel_ms(4)
                        g.close()
                        alg = B * (2.0 * P + 4.0 * N) + (8.0 * csr[1].size if nnz > 1 else 8.0 * N)
                        gbs = alg / (ms * 1e-3) / 1e9
                        rows.append(dict(height=H, width=W, nodes=N, csr=kind, nnz_per_row=nnz,
                                         kernel="fused (node-major rows)" if fused else ("k_project_ell1" if nnz == 1 else "k_project_csr"),
                                         batch=B, launches_sampled=ns, ms_per_launch=round(ms, 5),
                                         alg_bytes_per_launch=int(alg), gbs=round(gbs, 1), frac_of_hbm_peak=round(gbs / peak, 4)))
                        print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    print("wrote", args.out, len(rows), "points; hbm peak", peak)


if __name__ == "__main__":
    main()
