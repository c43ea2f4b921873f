"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): plain-Python restatement of the structured
model's seam detection, P3DModel_<FP>::identifyOverlap (cpp/lib/P3DModel.ipp:893-1127), its helpers
get_low_nidx (:1346-1354) / adjust_solution (:144-157) and the ray-caster triangle list extract_tris
(:234-317).  The reference's kd-tree range query (cpp/raycast/pspKdtree.c:225-256: squared distance in
double <= range^2) is an exhaustive numpy distance test here, so this stays a small-case checker.

Pinned by the reference's own unit test cpp/test/test_p3dmodel.cpp:166-240 (ExactNodeOverlap,
TolNodeOverlap, NodeIterator on the fixture of cpp/test/test_grid_utils.cpp:49-123), reproduced in
tests/test_p3d_model.py.
"""
import numpy as np


def zone_starts(sizes):
    return np.concatenate([[0], np.cumsum([j * k for j, k in sizes])]).astype(int)


def identify_overlap(xyz, sizes, tol):
    """xyz float32 [N,3]; sizes = [(J, K)] per zone.  Returns (overlap_pts dict, nonuniq, uniq)."""
    tol = np.float32(tol)
    tol = tol if tol > np.float32(1e-12) else np.float32(1e-12)
    rng2 = float(tol) * float(tol)
    starts = zone_starts(sizes)
    gidx = {}
    edge = []
    for zn, (J, K) in enumerate(sizes):
        for idx in range(J * K):
            r, c = divmod(idx, J)
            gidx[starts[zn] + idx] = (zn, c, r)
            if r == 0 or r == K - 1 or c == 0 or c == J - 1:
                edge.append(starts[zn] + idx)
    edge = np.array(edge, int)
    pos = xyz.astype(np.float32).astype(np.float64)[edge]
    pairs = set()            # the multimap, both directions
    oset, seen, uniq = set(), set(), set()
    for e, nidx in enumerate(edge):
        d = pos - pos[e]
        near = edge[(d * d).sum(1) <= rng2]
        zn, j, k = gidx[nidx]
        J, K = sizes[zn]
        others = set()
        for other in near:
            if other == nidx:
                continue
            oz, oj, ok = gidx[other]
            if oz == zn:
                wrapped = False
                if j == oj:
                    wrapped = (k == 0 and ok == K - 1) or (k == K - 1 and ok == 0)
                elif k == ok:
                    wrapped = (j == 0 and oj == J - 1) or (j == J - 1 and oj == 0)
                if not wrapped:
                    continue
            others.add(int(other))
        if not others:
            continue
        if nidx not in seen and nidx not in uniq:
            uniq.add(nidx)
            seen.add(nidx)
        for other in others:
            seen.add(other)
            lo, hi = min(nidx, other), max(nidx, other)
            if (lo, hi) not in pairs:
                oset.update((lo, hi))
                pairs.add((lo, hi))
                pairs.add((hi, lo))
    overlap_pts = {}
    for a, b in sorted(pairs):
        if a != b:
            overlap_pts.setdefault(int(a), []).append(int(b))
    return overlap_pts, len(oset), len(uniq)


def get_low_nidx(overlap_pts, nidx):
    v = overlap_pts.get(nidx)
    return v[0] if v and v[0] < nidx else nidx


def adjust_solution(overlap_pts, sol):
    for curr in sorted(overlap_pts):
        for alt in overlap_pts[curr]:
            if curr < alt:
                sol[alt] = sol[curr]
    return sol


def extract_tri_nodes(sizes):
    starts = zone_starts(sizes)
    out = []
    for zn, (J, K) in enumerate(sizes):
        for q in range((J - 1) * (K - 1)):
            klo, jlo = divmod(q, J - 1)
            i0 = starts[zn] + klo * J + jlo
            i1, i2, i3 = i0 + 1, i0 + 1 + J, i0 + J
            out += [i0, i1, i2, i2, i3, i0]
    return np.array(out, np.int32)
