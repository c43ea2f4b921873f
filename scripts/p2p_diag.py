"""2-GPU diagnostics: raw P2P copy bandwidth (torch), and the library's fused projection with peer
stores in ONE process (cudaDeviceEnablePeerAccess) to separate the IPC mapping from the kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

print(torch.cuda.device_count(), "gpus; can_access_peer 0->1:", torch.cuda.can_device_access_peer(0, 1))
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda:0")
y = torch.empty(1 << 28, dtype=torch.float32, device="cuda:1")
for _ in range(2):
    y.copy_(x)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
t = time.perf_counter()
for _ in range(5):
    y.copy_(x)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
print("torch P2P copy GB/s:", 5 * x.numel() * 4 / (time.perf_counter() - t) / 1e9)
del x, y

import upsp_b200 as up
synth = up.synth
H = W = 1024; N = 500_000; Fl = 2048; R = 2
FTOT = int(sys.argv[1]) if len(sys.argv) > 1 else Fl * R   # row length of the node-major buffers
frames = synth.make_frames_fast(64, H, W, seed=1)
packed = synth.pack_12bit(frames.reshape(64, -1))
csr = synth.make_projection(N, H, W, seed=1)
ctxs = []
for r in range(R):
    g = up.PspGpu(1, N, FTOT, device=r, rank=r, n_ranks=R)
    g.set_camera(0, W, H); g.set_projection(0, *csr)
    g.set_options(registration=up.REG_GIVEN)
    g.set_warp_matrices(0, 0, synth.make_warps(g.n_frames, seed=r))
    for o in range(0, Fl, 64):
        g.push_frames(0, packed, up.PIX_PACKED12, o, 64)
    ctxs.append(g)
up.connect_local(ctxs)
for it in range(3):
    for g in ctxs:
        g.reset_run(); g.set_kernel_sampling(1)
    t = time.perf_counter()
    for g in ctxs:
        g.process_frames(0, Fl)
    for g in ctxs:
        g.sync()
    dt = time.perf_counter() - t
    print(f"in-process 2-GPU fused phase 1: {dt*1e3:.1f} ms for {Fl} frames/GPU; k_project_fused mean",
          [round(g.kernel_ms(4)[0], 3) for g in ctxs], "ms per 128-frame batch")
print("row length", FTOT)
