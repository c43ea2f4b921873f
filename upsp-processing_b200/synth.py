"""Synthetic workloads for tests and bench: 12-bit cine-like frames, a
one-nearest-pixel-per-node projection matrix, fiducial patch clusters, P3D-style seam
overlaps, tunnel conditions.  Pure numpy; seeded; no reference code involved."""
from __future__ import annotations

import numpy as np


def make_frames(n_frames, height, width, seed=0, hot_frames=0.25, jitter=True, noise=8.0, texture=250.0):
    """u16 [F,H,W]: 1800 + 600 sin(x/37) cos(y/53) + 300 gauss blob + a fixed fine texture (sum of
    random sinusoids, wavelengths 5-40 px: paint speckle / model edges, what makes the image
    registration well posed) + N(0,noise), slow drift over the run (so the detrend has something
    to remove), optional sub-pixel translation jitter per frame (applied analytically), <= 5
    injected hot pixels (4095) in a fraction of the frames.
    Returns (frames, true_shift[F,2])."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float32)
    cx, cy = 0.55 * width, 0.45 * height
    K = 32
    wl = rng.uniform(5.0, 40.0, K)
    ang = rng.uniform(0, 2 * np.pi, K)
    kx, ky = 2 * np.pi / wl * np.cos(ang), 2 * np.pi / wl * np.sin(ang)
    ph = rng.uniform(0, 2 * np.pi, K)
    amp = texture * rng.uniform(0.5, 1.0, K) / np.sqrt(K)
    frames = np.empty((n_frames, height, width), np.uint16)
    shifts = np.zeros((n_frames, 2), np.float32)
    for f in range(n_frames):
        dx, dy = (rng.uniform(-1, 1, 2) if (jitter and f > 0) else (0.0, 0.0))
        shifts[f] = (dx, dy)
        xs, ys = x + dx, y + dy
        t = f / max(n_frames, 1)
        base = (1800.0 + 600.0 * np.sin(xs / 37.0) * np.cos(ys / 53.0)
                + 300.0 * np.exp(-(((xs - cx) / (0.2 * width)) ** 2 + ((ys - cy) / (0.2 * height)) ** 2)))
        if texture:
            tex = np.zeros_like(base)
            for k in range(K):
                tex += np.float32(amp[k]) * np.sin(np.float32(kx[k]) * xs + np.float32(ky[k]) * ys + np.float32(ph[k]))
            base = base + tex
        base *= (1.0 + 0.03 * np.sin(2.0 * np.pi * t) + 0.02 * t)
        img = base + rng.normal(0.0, noise, base.shape)
        frames[f] = np.clip(np.rint(img), 0, 4095).astype(np.uint16)
        if rng.random() < hot_frames:
            k = int(rng.integers(1, 6))
            frames[f].reshape(-1)[rng.integers(0, height * width, k)] = 4095
    return frames, shifts


def make_projection(n_nodes, height, width, kind="surface", seed=1, skipped_frac=0.02,
                    weights=False):
    """CSR (rowptr, col, val) with <= 1 entry per row, col = y*W+x, val = 1 (or a camera
    weight in (0,1]).  kind="surface": node order follows image raster order with local
    scatter (mesh-like locality, ~1 node per pixel neighbourhood); kind="random": uniformly
    random pixels (worst-case gather)."""
    rng = np.random.default_rng(seed)
    npix = height * width
    seen = rng.random(n_nodes) >= skipped_frac
    if kind == "random":
        cols = rng.integers(0, npix, n_nodes)
    else:
        base = np.sort(rng.integers(0, npix, n_nodes))
        dx = rng.integers(-3, 4, n_nodes)
        dy = rng.integers(-3, 4, n_nodes)
        yy = np.clip(base // width + dy, 0, height - 1)
        xx = np.clip(base % width + dx, 0, width - 1)
        cols = yy * width + xx
    rowptr = np.zeros(n_nodes + 1, np.int32)
    rowptr[1:] = np.cumsum(seen)
    col = cols[seen].astype(np.int32)
    val = (rng.uniform(0.2, 1.0, col.size).astype(np.float32) if weights
           else np.ones(col.size, np.float32))
    return rowptr, col, val


def make_multi_nnz_projection(n_nodes, height, width, nnz_per_row=4, seed=2):
    """cfg-5 variant: several weighted pixels per node (filter-folded stencil)."""
    rng = np.random.default_rng(seed)
    npix = height * width
    counts = rng.integers(0, nnz_per_row + 1, n_nodes)
    rowptr = np.zeros(n_nodes + 1, np.int32)
    rowptr[1:] = np.cumsum(counts)
    nnz = int(rowptr[-1])
    centre = np.repeat(rng.integers(0, npix, n_nodes), counts)
    col = np.clip(centre + rng.integers(-2, 3, nnz) + width * rng.integers(-2, 3, nnz), 0, npix - 1)
    val = rng.uniform(0.05, 0.6, nnz).astype(np.float32)
    return rowptr, col.astype(np.int32), val


def make_patches(height, width, n_targets=12, seed=3, bound_pts=2, buffer=1, half=3,
                 overlap_pair=False):
    """Single-target clusters laid out like get_target_boundary (cpp/lib/patches.ipp:290-327):
    interior = the target's bounding box, boundary = a `bound_pts`-thick frame `buffer` pixels
    outside it, both enumerated x-outer / y-inner and clipped to the frame.
    overlap_pair adds a cluster whose boundary frame crosses an earlier cluster's interior
    (exercises the in-order dependency between clusters).
    Returns (bounds, internal): lists over clusters of (x[], y[])."""
    rng = np.random.default_rng(seed)
    centres = []
    tries = 0
    margin = half + bound_pts + buffer + 2
    while len(centres) < n_targets and tries < 10000:
        tries += 1
        c = (int(rng.integers(margin, width - margin)), int(rng.integers(margin, height - margin)))
        if all(abs(c[0] - o[0]) > 4 * margin or abs(c[1] - o[1]) > 4 * margin for o in centres):
            centres.append(c)
    if overlap_pair and centres:
        c0 = centres[0]
        centres.append((min(c0[0] + half + buffer + 2, width - 2), c0[1]))
    # one target hugging the frame edge (clipping path)
    centres.append((1, height // 2))
    bounds, internal = [], []
    for (cx, cy) in centres:
        x0, x1, y0, y1 = cx - half, cx + half, cy - half, cy + half
        ix, iy, bx, by = [], [], [], []
        for x in range(x0, x1 + 1):
            for y in range(y0, y1 + 1):
                if 0 <= x < width and 0 <= y < height:
                    ix.append(x)
                    iy.append(y)
        o = bound_pts + buffer
        for x in range(x0 - o, x1 + o + 1):
            for y in range(y0 - o, y1 + o + 1):
                if (x < x0 - buffer or x > x1 + buffer or y < y0 - buffer or y > y1 + buffer):
                    if 0 <= x < width and 0 <= y < height:
                        bx.append(x)
                        by.append(y)
        bounds.append((np.array(bx, np.uint32), np.array(by, np.uint32)))
        internal.append((np.array(ix, np.uint32), np.array(iy, np.uint32)))
    return bounds, internal


def flatten_patches(bounds, internal):
    """-> (bounds_off, bx, by, internal_off, ix, iy) as upsp_gpu_set_patches wants them."""
    n = len(bounds)
    bo = np.zeros(n + 1, np.int32)
    io = np.zeros(n + 1, np.int32)
    for i in range(n):
        bo[i + 1] = bo[i] + len(bounds[i][0])
        io[i + 1] = io[i] + len(internal[i][0])
    cat = lambda l, k: (np.concatenate([a[k] for a in l]).astype(np.uint32) if n else np.zeros(0, np.uint32))
    return bo, cat(bounds, 0), cat(bounds, 1), io, cat(internal, 0), cat(internal, 1)


def make_overlap(n_nodes, n_groups=50, seed=4):
    """P3D seam overlap groups: {curr: [alt, ...]} as P3DModel_::overlap_pts_ holds them
    (symmetric, pairwise)."""
    rng = np.random.default_rng(seed)
    ov = {}
    for _ in range(n_groups):
        k = int(rng.integers(2, 4))
        grp = [int(v) for v in rng.choice(n_nodes, k, replace=False)]
        for a in grp:
            ov.setdefault(a, [])
            for b in grp:
                if b != a and b not in ov[a]:
                    ov[a].append(b)
    return ov


def overlap_src_index(n_nodes, overlap):
    """The array upsp_gpu_set_overlap_remap takes: out[n] = sol[src[n]] is what P3DModel_::adjust_solution
    (cpp/lib/P3DModel.ipp:144-157) does to a solution vector (lower node index of an overlap group wins)."""
    idx = np.arange(n_nodes, dtype=np.int32)
    for curr in sorted(overlap):
        for alt in overlap[curr]:
            if curr < alt:
                idx[alt] = idx[curr]
    return idx


def make_warps(n_frames, seed=5, trans=1.0, lin=5e-4):
    """Per-frame inverse affine maps near identity: translation within +-trans px, linear part
    within +-lin of identity.  f32 [F,6] row-major 2x3."""
    rng = np.random.default_rng(seed)
    m = np.zeros((n_frames, 2, 3), np.float32)
    m[:, 0, 0] = m[:, 1, 1] = 1.0
    m[:, :, :2] += rng.uniform(-lin, lin, (n_frames, 2, 2)).astype(np.float32)
    m[:, :, 2] = rng.uniform(-trans, trans, (n_frames, 2)).astype(np.float32)
    return m.reshape(n_frames, 6)


def tunnel_conditions(n_nodes, seed=6):
    """paint cal (a..f), qbar, ps, steady-state Cp [N], model temperature [N] (degF)."""
    rng = np.random.default_rng(seed)
    cal = np.array([0.62, -1.3e-3, 2.1e-6, 2.4e-4, 3.0e-7, -1.1e-9], np.float32)
    qbar, ps = np.float32(251.3), np.float32(1456.8)
    steady = rng.uniform(-1.2, 0.9, n_nodes).astype(np.float32)
    temp = (68.0 + rng.normal(0, 1.5, n_nodes)).astype(np.float32)
    return cal, qbar, ps, steady, temp


def pack_12bit(pix):
    """MSB-first 12-bit packing of an even number of pixels (the .mraw / packed .cine layout
    that unpack_12bit, cpp/lib/PSPVideo.cpp:134-150, decodes).  pix: u16 [..., npix]."""
    pix = np.asarray(pix, np.uint16)
    a, b = pix[..., 0::2], pix[..., 1::2]
    buf = np.empty(pix.shape[:-1] + (pix.shape[-1] * 3 // 2,), np.uint8)
    buf[..., 0::3] = a >> 4
    buf[..., 1::3] = ((a & 0x0F) << 4) | (b >> 8)
    buf[..., 2::3] = b & 0xFF
    return buf


def make_frames_fast(n_frames, height, width, seed=0, noise=8.0, hot_frames=0.25):
    """Bench-sized variant of make_frames: one base field, per-frame drift + integer noise."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float32)
    base = (1800.0 + 600.0 * np.sin(x / 37.0) * np.cos(y / 53.0)
            + 300.0 * np.exp(-(((x - 0.55 * width) / (0.2 * width)) ** 2
                               + ((y - 0.45 * height) / (0.2 * height)) ** 2))).astype(np.float32)
    frames = np.empty((n_frames, height, width), np.uint16)
    for f in range(n_frames):
        t = f / max(n_frames, 1)
        img = base * np.float32(1.0 + 0.03 * np.sin(2.0 * np.pi * t)) \
            + rng.standard_normal(base.shape, dtype=np.float32) * np.float32(noise)
        frames[f] = np.clip(np.rint(img), 0, 4095).astype(np.uint16)
        if rng.random() < hot_frames:
            frames[f].reshape(-1)[rng.integers(0, height * width, int(rng.integers(1, 6)))] = 4095
    return frames


def write_job(job_dir, *, frames, csr, fmt="p12", registration="none", interp="linear", warp=None,
              patches=None, remap=None, cal=None, qbar=1.0, ps=1.0, steady=None, model_temp=None,
              degree=6):
    """Write a psp_process_b200 job directory (host/psp_process_b200.cpp documents the layout).
    frames: list over cameras of u16 [F,H,W]; csr: list of (rowptr, col, val); patches: list of
    flatten_patches() tuples or None; warp: list of [F,6] f32 (registration = given)."""
    import os
    os.makedirs(job_dir, exist_ok=True)
    C = len(frames)
    F, H, W = frames[0].shape
    N = len(csr[0][0]) - 1
    with open(os.path.join(job_dir, "job.txt"), "w") as f:
        kv = dict(cameras=C, width=W, height=H, number_frames=F, msize=N, format=fmt,
                  registration=registration, pixel_interpolation=interp,
                  target_patcher="polynomial" if patches else "none", qbar=repr(float(qbar)),
                  ps=repr(float(ps)), degree=degree)
        for k, v in zip("abcdef", cal):
            kv["cal_" + k] = repr(float(v))
        for k, v in kv.items():
            f.write(f"{k} = {v}\n")
    for c in range(C):
        b = os.path.join(job_dir, f"cam{c}")
        fr = np.ascontiguousarray(frames[c], np.uint16)
        (pack_12bit(fr.reshape(F, -1)) if fmt == "p12" else fr).tofile(b + ".frames")
        np.asarray(csr[c][0], np.int32).tofile(b + ".rowptr")
        np.asarray(csr[c][1], np.int32).tofile(b + ".col")
        np.asarray(csr[c][2], np.float32).tofile(b + ".val")
        if registration == "given":
            np.asarray(warp[c], np.float32).tofile(b + ".warp")
        if registration == "pixel":
            fr[0].tofile(b + ".first")
        if patches:
            for name, arr, dt in zip(("boff", "bx", "by", "ioff", "ix", "iy"), patches[c],
                                     (np.int32, np.uint32, np.uint32, np.int32, np.uint32, np.uint32)):
                np.asarray(arr, dt).tofile(b + ".patch_" + name)
    if remap is not None:
        np.asarray(remap, np.int32).tofile(os.path.join(job_dir, "remap.i32"))
    np.asarray(steady, np.float32).tofile(os.path.join(job_dir, "steady.f32"))
    np.asarray(model_temp, np.float32).tofile(os.path.join(job_dir, "model_temp.f32"))


def make_sphere_mesh(n_lat=24, n_lon=48, radius=5.0, center=(0.0, 0.0, 0.0), bump=0.0, seed=0):
    """Closed triangulated sphere (optionally with a smooth radial bump so that it self-occludes a
    little): xyz [N,3] f32, outward unit vertex normals [N,3] f32, triangles [T,3] i32 of node indices
    (the reference's triNodes for an unstructured .tri grid)."""
    rng = np.random.default_rng(seed)
    th = np.linspace(0.0, np.pi, n_lat + 2)[1:-1]
    ph = np.linspace(0.0, 2 * np.pi, n_lon, endpoint=False)
    T, P = np.meshgrid(th, ph, indexing="ij")
    d = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1).reshape(-1, 3)
    d = np.concatenate([[[0, 0, 1.0]], d, [[0, 0, -1.0]]])
    r = radius * (1.0 + bump * np.sin(3 * np.arctan2(d[:, 1], d[:, 0])) * np.sin(2 * np.arccos(np.clip(d[:, 2], -1, 1))))
    xyz = (d * r[:, None] + np.asarray(center)[None, :] + rng.normal(0, 1e-4, d.shape)).astype(np.float32)
    tri = []
    idx = lambda i, j: 1 + i * n_lon + (j % n_lon)
    for j in range(n_lon):
        tri.append([0, idx(0, j), idx(0, j + 1)])
        tri.append([len(d) - 1, idx(n_lat - 1, j + 1), idx(n_lat - 1, j)])
    for i in range(n_lat - 1):
        for j in range(n_lon):
            tri.append([idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)])
            tri.append([idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)])
    return xyz, d.astype(np.float32), np.asarray(tri, np.int32)


def make_projection_scene(n_lat=40, n_lon=80, seed=0):
    """A model (bumpy sphere) with a small occluding sphere between it and the camera, a camera
    with lens distortion looking at it, and some non-data nodes: everything create_projection_mat
    consumes.  Returns dict(xyz, normals, tri, is_data, rvec, tvec, K, dist, width, height, thresh)."""
    rng = np.random.default_rng(seed)
    x1, n1, t1 = make_sphere_mesh(n_lat, n_lon, 5.0, (0, 0, 0), bump=0.12, seed=seed)
    x2, n2, t2 = make_sphere_mesh(8, 16, 0.9, (1.5, 1.0, -8.0), seed=seed + 1)
    xyz = np.concatenate([x1, x2])
    normals = np.concatenate([n1, n2])
    tri = np.concatenate([t1, t2 + len(x1)])
    is_data = np.ones(len(xyz), np.uint8)
    is_data[rng.choice(len(xyz), len(xyz) // 50, replace=False)] = 0
    rvec = np.array([0.05, -0.08, 0.3])
    tvec = np.array([0.3, -0.2, 30.0])
    K = np.array([[2400.0, 0, 255.3], [0, 2390.0, 250.8], [0, 0, 1]])
    dist = np.array([-0.12, 0.06, 0.001, -0.0007, 0.01])
    return dict(xyz=xyz, normals=normals, tri=tri, is_data=is_data, rvec=rvec, tvec=tvec, K=K, dist=dist,
                width=512, height=512, thresh=float(np.deg2rad(100.0)))
