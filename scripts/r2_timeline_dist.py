"""Per-kernel timeline of one process_frames call on every rank of a torchrun job (see r2_timeline.py)."""
import os, sys, time, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import upsp_b200 as up
from upsp_b200 import synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rank, world, local, dist = bench.dist_setup(int(os.environ.get("WORLD_SIZE", "1")))
args = types.SimpleNamespace(height=1024, width=1024, nodes=500000, frames=F, targets=32, distinct=128, degree=6,
                             batch=0, csr="surface", registration="given", cams=1, overlap_groups=0, config=1, exchange="peer")
wl = bench.build_workload(args, synth)
g = bench.configure(up, wl, args, rank, world, local, 0, dist)
for o in range(0, g.n_frames, 128):
    n = min(128, g.n_frames - o)
    g.push_frames(0, wl["packed"][0][:n], up.PIX_PACKED12, o, n)
g.sync()
for _ in range(2):
    g.reset_run(); g.process_frames(0, g.n_frames); g.sync(); bench.barrier(dist)
g.reset_run()
bench.barrier(dist)
t0 = time.perf_counter(); g.process_frames(0, g.n_frames); g.sync(); t_plain = (time.perf_counter() - t0) * 1e3
bench.barrier(dist)
g.reset_run()
g.timeline(True)
bench.barrier(dist)
t0 = time.perf_counter(); g.process_frames(0, g.n_frames); g.sync(); t_tl = (time.perf_counter() - t0) * 1e3
rec = np.asarray(g.timeline_read(8192)).reshape(-1, 3)
g.timeline(False)
bench.barrier(dist)
names = {0: "scan", 3: "patch", 4: "project"}
out = [f"rank {rank}/{world} frames={F} plain {t_plain:.2f} ms, with timeline {t_tl:.2f} ms"]
for cls in sorted(set(rec[:, 0].astype(int))):
    r = rec[rec[:, 0] == cls]
    d = r[:, 2] - r[:, 1]
    out.append(f"  {names.get(cls, cls):8s} n={len(r):4d} mean {d.mean():.4f} ms  min {d.min():.4f} max {d.max():.4f}")
prj = rec[rec[:, 0] == 4][:, 1:]
gaps = prj[1:, 0] - prj[:-1, 1]
out.append(f"  project-to-project gaps: mean {gaps.mean():.4f} ms max {gaps.max():.4f}; first 6 project intervals: " +
           ", ".join(f"{a:.3f}->{b:.3f}" for a, b in prj[:6]))
for i in range(world):
    bench.barrier(dist)
    if i == rank:
        print("\n".join(out), flush=True)
