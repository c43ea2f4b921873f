"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): float64 / cv2 model of the target handling in
InitializeImagePatches (cpp/exec/psp_process.cpp:2095-2123): getTargets (:56-114), map_points_to_image
(cv::projectPoints, cpp/lib/CameraCal.ipp:225-239) and get_target_diameters (:116-165, get_perpendicular
cpp/utils/cv_extras.ipp:30-66).  The image positions come from cv2.projectPoints itself (the OpenCV entry point the
reference calls); the ray cast is an exhaustive Moeller-Trumbore in float64, so visibility decisions are only
comparable away from their thresholds and diameters to float precision.  Parity unpinned by the reference (it has no
test for these functions).
"""
import numpy as np


def project(cv2, xyz, rvec, tvec, K, dist):
    pts = np.asarray(xyz, np.float32).reshape(-1, 1, 3)
    uv, _ = cv2.projectPoints(pts, np.asarray(rvec, float), np.asarray(tvec, float), np.asarray(K, float), np.asarray(dist, float))
    return uv.reshape(-1, 2).astype(np.float32)


def nearest_hit(orig, d, verts, tris):
    a, b, c = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    e1, e2 = b - a, c - a
    p = np.cross(d, e2)
    det = (e1 * p).sum(1)
    ok = np.abs(det) > 1e-14
    inv = np.where(ok, 1.0 / np.where(ok, det, 1), 0)
    s = orig - a
    u = (s * p).sum(1) * inv
    q = np.cross(s, e1)
    v = (q * d).sum(1) * inv
    t = (e2 * q).sum(1) * inv
    hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 0)
    return t[hit].min() if hit.any() else None


def perpendicular(n):
    n = n / np.linalg.norm(n)
    m = int(np.argmax(np.abs(n))) if not (abs(n[0]) > abs(n[1]) and abs(n[0]) > abs(n[2])) else 0
    if m == 0:
        out = np.array([-(n[1]) / n[0], 1.0, 0.0])
    elif m == 1:
        out = np.array([1.0, -(n[0]) / n[1], 0.0])
    else:
        out = np.array([1.0, 0.0, -(n[0]) / n[2]])
    return out / np.linalg.norm(out)


def visible_targets(cv2, targets, xyz, normals, tris, rvec, tvec, K, dist, width, height, oblique_angle, diam_sf, node_normals=None):
    """targets: [(x, y, z, diameter)] float32.  Returns [(index, u, v, diameter_px, margin)]; margin = how far the
    closest decision (frame edge in px, occlusion distance, obliqueness in rad) is from flipping."""
    R = cv2.Rodrigues(np.asarray(rvec, float))[0]
    center = (-R.T @ np.asarray(tvec, float)).astype(np.float32).astype(np.float64)
    verts = np.asarray(xyz, np.float32).astype(np.float64)
    nrm = np.asarray(normals, np.float32).astype(np.float64)                 # Model::get_n(): the obliqueness test of getTargets
    nrm_d = nrm if node_normals is None else np.asarray(node_normals, np.float32).astype(np.float64)   # Node::get_normal(): diameters
    thresh = np.deg2rad(180.0 - min(oblique_angle + 5.0, 90.0))
    out = []
    for i, (x, y, z, diam) in enumerate(targets):
        pos = np.array([x, y, z], np.float32).astype(np.float64)
        u, v = project(cv2, pos, rvec, tvec, K, dist)[0]
        if u < 0 or v < 0 or u >= width or v >= height:
            continue
        d = pos - center
        dist_eye = np.linalg.norm(d)
        d = d / dist_eye
        t = nearest_hit(center, d, verts, tris)
        if t is None or t < dist_eye - 1e-3:
            continue
        hit = center + t * d
        nn = int(np.argmin(((verts - hit) ** 2).sum(1)))
        ang = np.arccos(np.clip(nrm[nn] @ d, -1, 1))
        if not ang > thresh:
            continue
        n_t = nrm_d[int(np.argmin(((verts - pos) ** 2).sum(1)))]
        a = perpendicular(n_t)
        b = np.cross(a, n_t)
        total = 0.0
        for j in range(4):
            th = j * np.pi / 2
            est = pos + 0.5 * diam * np.cos(th) * a + 0.5 * diam * np.sin(th) * b
            pu, pv = project(cv2, est, rvec, tvec, K, dist)[0]
            total += 2.0 * np.hypot(float(pu) - float(u), float(pv) - float(v))
        margin = min(u, v, width - u, height - v)
        out.append((i, np.float32(u), np.float32(v), total / 4.0 * diam_sf, float(min(margin, ang - thresh))))
    return out
