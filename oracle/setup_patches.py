"""CPU restatement (TEST INFRASTRUCTURE ONLY) of the reference's fiducial patch geometry, in plain
Python loops for small cases:
    cluster_points          cpp/lib/patches.ipp:240-275
    get_target_boundary     cpp/lib/patches.ipp:279-326
    get_cluster_boundary    cpp/lib/patches.ipp:330-487
    PatchClusters ctor      cpp/lib/patches.ipp:15-54
    threshold_bounds        cpp/lib/patches.ipp:59-94
Parity unpinned: the reference has no test or fixture for these functions (cpp/test has none for
patches), so this restatement and host/patch_geometry.hpp are two independent readings of the same
source that the tests hold against each other."""
import math

import numpy as np


def cluster_points(targs, bound_pts=4):
    """targs: list of (u, v, diameter) float32 triples -> list of clusters (lists of triples)"""
    targs = [tuple(np.float32(x) for x in t) for t in targs]
    pts = list(range(len(targs)))
    clusters = []
    while pts:
        cur = [targs[pts.pop(0)]]
        queue = [0]
        while queue:
            ref = cur[queue.pop(0)]
            keep = []
            for idx in pts:
                o = targs[idx]
                dx, dy = float(np.float32(ref[0] - o[0])), float(np.float32(ref[1] - o[1]))
                if math.sqrt(dx * dx + dy * dy) <= float(np.float32(bound_pts)) + 0.5 * float(np.float32(ref[2] + o[2])):
                    queue.append(len(cur))
                    cur.append(o)
                else:
                    keep.append(idx)
            pts = keep
        clusters.append(cur)
    return clusters


def _limits(t):
    u, v, d = (float(x) for x in t)
    return (math.floor(u - 0.5 * d), math.floor(v - 0.5 * d)), (math.ceil(u + 0.5 * d), math.ceil(v + 0.5 * d))


def get_target_boundary(t, bound_pts=2, buffer=0):
    (x0, y0), (x1, y1) = _limits(t)
    internal = [(x, y) for x in range(x0, x1 + 1) for y in range(y0, y1 + 1)]
    bounds = [(x, y) for x in range(x0 - bound_pts - buffer, x1 + bound_pts + buffer + 1)
              for y in range(y0 - bound_pts - buffer, y1 + bound_pts + buffer + 1)
              if x < x0 - buffer or x > x1 + buffer or y < y0 - buffer or y > y1 + buffer]
    return internal, bounds


def get_cluster_boundary(targs, bound_pts=2, buffer=0):
    lims = [_limits(t) for t in targs]
    pad = bound_pts + buffer
    tx0 = min(l[0][0] for l in lims) - pad
    ty0 = min(l[0][1] for l in lims) - pad
    tx1 = max(max(l[1][0] for l in lims), 0) + pad       # t_max starts at (0, 0) in the reference
    ty1 = max(max(l[1][1] for l in lims), 0) + pad
    dx, dy = tx1 - tx0 + 1, ty1 - ty0 + 1
    g = np.zeros((dx, dy), np.int64)
    for (a, b) in lims:
        g[a[0] - tx0:b[0] - tx0 + 1, a[1] - ty0:b[1] - ty0 + 1] = 2
    for x in range(dx):
        ys = np.flatnonzero(g[x] == 2)
        if len(ys):
            g[x, ys[0]:ys[-1] + 1] = 2
    for y in range(dy):
        xs = np.flatnonzero(g[:, y] == 2)
        if len(xs):
            g[xs[0]:xs[-1] + 1, y] = 2
    internal, bounds = [], []
    for x in range(dx):
        mx = 0 if x <= pad else x - pad
        lx = min(x + pad, dx - 1) - mx + 1
        bmx = 0 if x <= buffer else x - buffer
        blx = min(x + buffer, dx - 1) - bmx + 1
        for y in range(dy):
            my = 0 if y <= pad else y - pad
            ly = min(y + pad, dy - 1) - my + 1
            if g[x, y] == 2:
                internal.append((x + tx0, y + ty0))
                continue
            if bound_pts > 0 and buffer > 0:
                bmy = 0 if y <= buffer else y - buffer
                bly = min(y + buffer, dy - 1) - bmy + 1
                if g[bmx:bmx + blx, bmy:bmy + bly].max() != 2 and g[mx:mx + lx, my:my + ly].max() == 2:
                    bounds.append((x + tx0, y + ty0))
                    g[x, y] = 1
                continue
            if bound_pts > 0 and g[mx:mx + lx, my:my + ly].max() == 2:
                bounds.append((x + tx0, y + ty0))
                g[x, y] = 1
    return internal, bounds


def patch_clusters(clusters, width, height, boundary_thickness, buffer_thickness, ref=None, thresh=0, offset=2):
    """-> list over clusters of (bounds [(x, y)], internal [(x, y)]) as PatchClusters holds them"""
    out = []
    for cl in clusters:
        if len(cl) > 1:
            internal, bounds = get_cluster_boundary(cl, boundary_thickness, buffer_thickness)
        else:
            internal, bounds = get_target_boundary(cl[0], boundary_thickness, buffer_thickness)
        inside = lambda p: 0 <= p[0] < width and 0 <= p[1] < height
        internal = [p for p in internal if inside(p)]
        bounds = [p for p in bounds if inside(p)]
        if ref is not None:
            kept = []
            for (x, y) in bounds:
                y0, x0 = max(0, y - offset), max(0, x - offset)
                w = min(width - 1, x + offset) - x0 + 1
                h = min(height - 1, y + offset) - y0 + 1
                if not (float(ref[y0:y0 + h, x0:x0 + w].min()) < thresh):
                    kept.append((x, y))
            bounds = kept
        out.append((bounds, internal))
    return out
