"""Timeline of one process_frames call: every kernel bracketed by CUDA events on its own stream, pipeline kept.
usage: python scripts/r2_timeline.py [frames]   (env: UPSP_PROJ, UPSP_PROJ_PAD, UPSP_DECODE_BPSM ...)"""
import os, sys, time, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import upsp_b200 as up
from upsp_b200 import synth
up.build.build()
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
args = types.SimpleNamespace(height=1024, width=1024, nodes=500000, frames=F, targets=32, distinct=128, degree=6,
                             batch=0, csr="surface", registration="given", cams=1, overlap_groups=0, config=1, exchange="peer")
wl = bench.build_workload(args, synth)
g = bench.configure(up, wl, args, 0, 1, 0, 0, None)
for o in range(0, g.n_frames, 128):
    n = min(128, g.n_frames - o)
    g.push_frames(0, wl["packed"][0][:n], up.PIX_PACKED12, o, n)
g.sync()
for _ in range(2):
    g.reset_run(); g.process_frames(0, g.n_frames); g.sync()
g.reset_run()
t0 = time.perf_counter(); g.process_frames(0, g.n_frames); g.sync(); t_plain = (time.perf_counter() - t0) * 1e3
g.reset_run()
g.timeline(True)
t0 = time.perf_counter(); g.process_frames(0, g.n_frames); g.sync(); t_tl = (time.perf_counter() - t0) * 1e3
rec = g.timeline_read(8192)
g.timeline(False)
rec = np.asarray(rec).reshape(-1, 3)
names = {0: "decode", 2: "warp", 3: "patch", 4: "project"}
print(f"mode={g.projection_mode() if hasattr(g,'projection_mode') else '?'} frames={F} plain {t_plain:.2f} ms, with timeline {t_tl:.2f} ms, records {len(rec)}")
for cls in sorted(set(rec[:, 0].astype(int))):
    r = rec[rec[:, 0] == cls]
    d = r[:, 2] - r[:, 1]
    print(f"  {names.get(cls, cls):8s} n={len(r):4d} mean {d.mean():.4f} ms  sum {d.sum():.2f} ms  first [{r[0,1]:.3f},{r[0,2]:.3f}] ")
# overlap: time during which decode and project intervals intersect
dec = rec[rec[:, 0] == 0][:, 1:]; prj = rec[rec[:, 0] == 4][:, 1:]
ov = 0.0
for a, b in dec:
    lo = np.maximum(prj[:, 0], a); hi = np.minimum(prj[:, 1], b)
    ov += np.clip(hi - lo, 0, None).sum()
print(f"  decode/project overlap {ov:.2f} ms; span {rec[:,2].max() - rec[:,1].min():.2f} ms")
for i in range(min(12, len(rec))):
    print("   ", names.get(int(rec[i, 0]), rec[i, 0]), f"{rec[i,1]:.3f} -> {rec[i,2]:.3f}")
