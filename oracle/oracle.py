"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes front-end for ``oracle/libupsp_oracle.so`` (the plain-C restatement of the
reference's psp_process frame chain, see ``upsp_oracle.c``) plus the pieces that
lean on OpenCV through ``cv2`` (the same library entry points the reference
calls: ``cv::findTransformECC`` / ``cv::warpAffine`` at
cpp/lib/registration.cpp:64-72).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` arm may import this module, and only as the checker.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libupsp_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("upsp_oracle.c", "upsp_oracle_setup.c")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def build_ref(ref_root: str = "/root/reference"):
    """Compile the reference's dependency-free post-processing tools from the reference tree into oracle/_ref/ (only where
    that tree exists, i.e. in the build container; the built files travel to the GPU box).  Returns the directory or None."""
    out = os.path.join(_HERE, "_ref")
    if os.path.isdir(os.path.join(ref_root, "cpp", "exec")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "REF=" + ref_root])
    return out if os.path.exists(os.path.join(out, "xyz_scalar_to_tbl")) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_get_gain.restype = C.c_float
        _lib.orc_get_gain.argtypes = [C.c_void_p, C.c_float, C.c_float]
        _lib.orc_fix_hot_pixels.restype = C.c_int
        _lib.orc_colpiv_qr_f32.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ---------------------------------------------------------------- a1 decode
def unpack_12bit(packed: np.ndarray) -> np.ndarray:
    packed = _c(packed, np.uint8).ravel()
    out = np.zeros(packed.size * 2 // 3, np.uint16)
    lib().orc_unpack_12bit(_p(packed), C.c_size_t(packed.size), _p(out))
    return out


def unpack_12bit_frames(packed: np.ndarray) -> np.ndarray:
    """[F, frame_bytes] packed -> [F, npix] u16, frames decoded in parallel (one frame per thread at a time)."""
    packed = _c(packed, np.uint8)
    F, fb = packed.shape
    out = np.zeros((F, fb // 3 * 2), np.uint16)
    lib().orc_unpack_12bit_frames(_p(packed), C.c_size_t(fb), C.c_int(F), _p(out))
    return out


def unpack_10bit(packed: np.ndarray, lut: np.ndarray | None = None) -> np.ndarray:
    packed = _c(packed, np.uint8).ravel()
    out = np.zeros(packed.size * 4 // 5, np.uint16)
    lutp = _p(_c(lut, np.uint16)) if lut is not None else None
    lib().orc_unpack_10bit(_p(packed), C.c_size_t(packed.size), _p(out), lutp)
    return out


def pack_12bit(pix: np.ndarray) -> np.ndarray:
    """Inverse of unpack_12bit (python/upsp/video/util.py:39-51 layout)."""
    pix = np.asarray(pix, np.uint16).ravel()
    buf = np.zeros(pix.size * 3 // 2, np.uint8)
    buf[0::3] = pix[0::2] >> 4
    buf[1::3] = ((pix[0::2] & 0x0F) << 4) | (pix[1::2] >> 8)
    buf[2::3] = pix[1::2] & 0xFF
    return buf


def pack_10bit(pix: np.ndarray) -> np.ndarray:
    pix = np.asarray(pix, np.uint16).ravel()
    a, b, c, d = pix[0::4], pix[1::4], pix[2::4], pix[3::4]
    buf = np.zeros(pix.size * 5 // 4, np.uint8)
    buf[0::5] = a >> 2
    buf[1::5] = ((a & 0x3) << 6) | (b >> 4)
    buf[2::5] = ((b & 0xF) << 4) | (c >> 6)
    buf[3::5] = ((c & 0x3F) << 2) | (d >> 8)
    buf[4::5] = d & 0xFF
    return buf


# ---------------------------------------------------------------- a2 hot px
def fix_hot_pixels(img: np.ndarray, thresh=4064, min_change=512, max_hot=5):
    out = _c(img, np.uint16).copy()
    n = lib().orc_fix_hot_pixels(_p(out), out.shape[0], out.shape[1], thresh, min_change, max_hot)
    return out, n


# ---------------------------------------------------------------- a3 warp
def warp_affine(img: np.ndarray, M: np.ndarray, interp: int = 1) -> np.ndarray:
    """cv::warpAffine(img, M, size, interp | WARP_INVERSE_MAP), BORDER_CONSTANT 0."""
    M = _c(M, np.float32).reshape(6)
    h, w = img.shape
    if img.dtype == np.uint16:
        src = _c(img, np.uint16)
        dst = np.empty_like(src)
        lib().orc_warp_affine_u16(_p(src), w, h, _p(M), interp, _p(dst), w, h)
    else:
        src = _c(img, np.float32)
        dst = np.empty_like(src)
        lib().orc_warp_affine_f32(_p(src), w, h, _p(M), interp, _p(dst), w, h)
    return dst


def ecc_cv2(ref32: np.ndarray, inp_u16: np.ndarray, max_iters=50, eps=1e-3):
    """cv::findTransformECC exactly as cpp/lib/registration.cpp:43-64 calls it
    (5-argument overload: no mask, gaussFiltSize 5).  Returns (M[2,3] f32, rho)."""
    import cv2

    inp = inp_u16.astype(np.float32)
    M = np.eye(2, 3, dtype=np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, max_iters, eps)
    rho, M = cv2.findTransformECC(ref32, inp, M, cv2.MOTION_AFFINE, crit, None, 5)
    return M, rho


def register_pixel_cv2(ref32, inp_u16, interp=1, max_iters=50, eps=1e-3):
    """upsp::register_pixel (cpp/lib/registration.cpp:32-81) through cv2."""
    import cv2

    M, _ = ecc_cv2(ref32, inp_u16, max_iters, eps)
    flag = cv2.INTER_LINEAR if interp == 1 else cv2.INTER_NEAREST
    out = cv2.warpAffine(inp_u16, M, (inp_u16.shape[1], inp_u16.shape[0]),
                         flags=flag | cv2.WARP_INVERSE_MAP)
    return out, M


# ---------------------------------------------------------------- QR / patch
def colpiv_qr_solve(A: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Eigen ColPivHouseholderQR<float>(A).solve(b) restated (scalar order)."""
    A = np.asfortranarray(A, dtype=np.float32).copy(order="F")
    rows, cols = A.shape
    h = np.zeros(cols, np.float32)
    t = np.zeros(cols, np.int32)
    nz = lib().orc_colpiv_qr_f32(_p(A), rows, cols, _p(h), _p(t))
    x = np.zeros(cols, np.float32)
    c = np.zeros(rows, np.float32)
    b = _c(b, np.float32)
    lib().orc_colpiv_qr_solve_f32(_p(A), rows, cols, _p(h), _p(t), nz, _p(b), _p(x), _p(c))
    return x


class Patches:
    """Flattened PatchClusters geometry (cpp/include/patches.h:74-90): per cluster the
    boundary pixel list and the interior pixel list."""

    def __init__(self, bounds, internal):
        # bounds/internal: list over clusters of (x[], y[]) integer arrays
        self.n = len(bounds)
        self.bounds_off = np.zeros(self.n + 1, np.int32)
        self.internal_off = np.zeros(self.n + 1, np.int32)
        for i in range(self.n):
            self.bounds_off[i + 1] = self.bounds_off[i] + len(bounds[i][0])
            self.internal_off[i + 1] = self.internal_off[i] + len(internal[i][0])
        cat = lambda l, k: (np.concatenate([np.asarray(a[k], np.uint32) for a in l])
                            if l else np.zeros(0, np.uint32))
        self.bx, self.by = _c(cat(bounds, 0), np.uint32), _c(cat(bounds, 1), np.uint32)
        self.ix, self.iy = _c(cat(internal, 0), np.uint32), _c(cat(internal, 1), np.uint32)


def patch_apply(img32: np.ndarray, p: Patches) -> np.ndarray:
    out = _c(img32, np.float32).copy()
    if p is not None and p.n:
        lib().orc_patch_apply(_p(out), out.shape[1], p.n, _p(p.bounds_off), _p(p.bx), _p(p.by),
                              _p(p.internal_off), _p(p.ix), _p(p.iy))
    return out


# ---------------------------------------------------------------- a5 filter
def gauss_fixed_taps(ksize: int) -> np.ndarray:
    """OpenCV's 16.16 fixed-point taps of GaussianBlur(CV_16U, (k,k), sigma=0): the bit-exact double kernel
    (getGaussianKernelBitExact, read through cv2.getGaussianKernel) rounded with error diffusion, the centre tap
    takes what is left of 65536 (getGaussianKernelFixedPoint_ED in OpenCV's smooth.dispatch.cpp)."""
    import cv2
    kern = cv2.getGaussianKernel(ksize, 0, cv2.CV_64F).ravel()
    half, err, total = [], 0.0, 0
    for i in range(ksize // 2):
        adj = kern[i] * 65536.0 + err
        v0 = int(np.rint(adj))
        err = adj - v0
        half.append(v0)
        total += v0
    return np.array(half + [65536 - 2 * total] + half[::-1], np.int32)


def ensure_gauss_taps(kind: int, ksize: int):
    """sizes beyond the three built-in small kernels need their taps registered (needs cv2)"""
    if kind == 1 and ksize > 7:
        q = gauss_fixed_taps(ksize)
        lib().orc_set_gauss_taps(ksize, _p(q))


def spatial_filter(img, kind, ksize):
    """cv::GaussianBlur(img,(k,k),0) (kind 1) / cv::blur(img,(k,k)) (kind 2) on u16 or f32."""
    ensure_gauss_taps(kind, ksize)
    h, w = img.shape
    if img.dtype == np.uint16:
        src = _c(img, np.uint16)
        dst = np.empty_like(src)
        rc = lib().orc_filter_u16(_p(src), _p(dst), w, h, kind, ksize)
    else:
        src = _c(img, np.float32)
        dst = np.empty_like(src)
        rc = lib().orc_filter_f32(_p(src), _p(dst), w, h, kind, ksize)
    if rc:
        raise ValueError("gaussian filter sizes other than 3, 5, 7 are not restated")
    return dst


# ---------------------------------------------------------------- a6 project
def project_frame(rowptr, col, val, frame32) -> np.ndarray:
    rowptr, col, val = _c(rowptr, np.int32), _c(col, np.int32), _c(val, np.float32)
    n = rowptr.size - 1
    out = np.zeros(n, np.float32)
    fr = _c(frame32, np.float32).ravel()
    lib().orc_project_frame(_p(rowptr), _p(col), _p(val), n, _p(fr), _p(out))
    return out


def identify_skipped(rowptrs) -> np.ndarray:
    """upsp::identify_skipped_nodes cpp/lib/projection.ipp:858-880."""
    n = rowptrs[0].size - 1
    any_nz = np.zeros(n, bool)
    for rp in rowptrs:
        any_nz |= (np.diff(rp) > 0)
    return np.nonzero(~any_nz)[0].astype(np.int32)


def overlap_remap(n_nodes: int, overlap: dict[int, list[int]]) -> np.ndarray:
    """Static form of P3DModel_::adjust_solution (cpp/lib/P3DModel.ipp:144-157): run the
    reference loop once on an index vector; out[n] = in[src_index[n]] afterwards."""
    idx = np.arange(n_nodes, dtype=np.int32)
    for curr in sorted(overlap):
        for alt in overlap[curr]:
            if curr < alt:
                idx[alt] = idx[curr]
    return idx


def apportion(value: int, n_bins: int):
    s = np.zeros(n_bins, np.int32)
    e = np.zeros(n_bins, np.int32)
    lib().orc_apportion(value, n_bins, _p(s), _p(e))
    return s, e


# ---------------------------------------------------------------- phase 1
class _P1Args(C.Structure):
    _fields_ = [
        ("n_cams", C.c_int), ("n_nodes", C.c_int), ("n_frames", C.c_int), ("first_frame", C.c_int),
        ("width", C.c_void_p), ("height", C.c_void_p), ("frames", C.c_void_p), ("warp", C.c_void_p),
        ("interp", C.c_int), ("hot_pixel_fix", C.c_int),
        ("rowptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p),
        ("n_clusters", C.c_void_p), ("bounds_off", C.c_void_p), ("bx", C.c_void_p),
        ("by", C.c_void_p), ("internal_off", C.c_void_p), ("ix", C.c_void_p), ("iy", C.c_void_p),
        ("n_skipped", C.c_int), ("skipped", C.c_void_p), ("remap", C.c_void_p),
        ("filter_kind", C.c_int), ("filter_size", C.c_int),
    ]


def _ptr_array(arrs):
    a = (C.c_void_p * len(arrs))()
    for i, x in enumerate(arrs):
        a[i] = x.ctypes.data if x is not None else None
    return a


def phase1(frames, csr, *, first_frame=0, warp=None, interp=1, patches=None, remap=None,
           hot_pixel_fix=True, sum_=None, sumsq=None, filter_kind=0, filter_size=0):
    """The frame loop of cpp/exec/psp_process.cpp:1743-1851 for one rank's slice.

    frames : list over cameras of u16 [F, H, W]
    csr    : list over cameras of (rowptr, col, val)
    warp   : list over cameras of f32 [F, 6] (or None) -- registration result
    Returns (intensity [F, N] f32, sum [N] f64, sumsq [N] f64).
    """
    nc = len(frames)
    frames = [_c(f, np.uint16) for f in frames]
    F = frames[0].shape[0]
    rowptr = [_c(c[0], np.int32) for c in csr]
    col = [_c(c[1], np.int32) for c in csr]
    val = [_c(c[2], np.float32) for c in csr]
    N = rowptr[0].size - 1
    width = np.array([f.shape[2] for f in frames], np.int32)
    height = np.array([f.shape[1] for f in frames], np.int32)
    skipped = identify_skipped(rowptr)
    keep = [frames, rowptr, col, val, width, height, skipped]
    a = _P1Args()
    a.n_cams, a.n_nodes, a.n_frames, a.first_frame = nc, N, F, first_frame
    a.width, a.height = width.ctypes.data, height.ctypes.data
    fa = _ptr_array(frames)
    a.frames = C.cast(fa, C.c_void_p)
    if warp is not None:
        warp = [(_c(w, np.float32).reshape(F, 6) if w is not None else None) for w in warp]
        wa = _ptr_array(warp)
        a.warp = C.cast(wa, C.c_void_p)
        keep += [warp, wa]
    a.interp, a.hot_pixel_fix = interp, int(hot_pixel_fix)
    ra, ca, va = _ptr_array(rowptr), _ptr_array(col), _ptr_array(val)
    a.rowptr, a.col, a.val = (C.cast(x, C.c_void_p) for x in (ra, ca, va))
    if patches is not None:
        ncl = np.array([(p.n if p is not None else 0) for p in patches], np.int32)
        empty_i = np.zeros(1, np.int32)
        empty_u = np.zeros(1, np.uint32)
        g = lambda p, n, e: (getattr(p, n) if p is not None else e)
        arrs = [_ptr_array([g(p, n, empty_i if n.endswith("off") else empty_u) for p in patches])
                for n in ("bounds_off", "bx", "by", "internal_off", "ix", "iy")]
        a.n_clusters = ncl.ctypes.data
        (a.bounds_off, a.bx, a.by, a.internal_off, a.ix, a.iy) = (C.cast(x, C.c_void_p) for x in arrs)
        keep += [ncl, arrs, empty_i, empty_u]
    a.n_skipped, a.skipped = skipped.size, skipped.ctypes.data
    a.filter_kind, a.filter_size = int(filter_kind), int(filter_size)
    ensure_gauss_taps(a.filter_kind, a.filter_size)
    if remap is not None:
        remap = _c(remap, np.int32)
        a.remap = remap.ctypes.data
    inten = np.empty((F, N), np.float32)
    if sum_ is None:
        sum_ = np.zeros(N, np.float64)
        sumsq = np.zeros(N, np.float64)
    lib().orc_phase1(C.byref(a), _p(inten), _p(sum_), _p(sumsq))
    del keep
    return inten, sum_, sumsq


def phase1_finals(sum_, sumsq, n_frames_total, remap=None):
    N = sum_.size
    avg = np.zeros(N, np.float32)
    rms = np.zeros(N, np.float32)
    rp = _p(_c(remap, np.int32)) if remap is not None else None
    lib().orc_phase1_finals(_p(sum_), _p(sumsq), N, C.c_uint(n_frames_total), rp, _p(avg), _p(rms))
    return avg, rms


def coverage(csr, remap=None) -> np.ndarray:
    """coverage = sum_c project(ones) (cpp/exec/psp_process.cpp:1953-1976), then
    adjust_solution for P3D."""
    cov = None
    for rowptr, col, val in csr:
        ncol = int(col.max()) + 1 if len(col) else 1
        c = project_frame(rowptr, col, val, np.ones(ncol, np.float32))
        cov = c if cov is None else (cov + c).astype(np.float32)
    if remap is not None:
        cov = cov[remap]
    return cov


# ---------------------------------------------------------------- transpose
def global_transpose(src_slices, n_nodes, n_frames):
    """global_transpose (cpp/exec/psp_process.cpp:707-771) with ranks simulated.
    src_slices: list over ranks of [F_r, N] f32.  Returns list over ranks of [N_s, F]."""
    R = len(src_slices)
    src = [_c(s, np.float32) for s in src_slices]
    ns, ne = apportion(n_nodes, R)
    dst = [np.zeros((int(ne[s]), n_frames), np.float32) for s in range(R)]
    sa, da = _ptr_array(src), _ptr_array(dst)
    lib().orc_global_transpose(sa, da, R, n_nodes, n_frames)
    return dst


# ---------------------------------------------------------------- phase 2
def transpoly_fit(data: np.ndarray, degree=6):
    """TransPolyFitter<float>(F, degree, 1).eval_fit(data, 1, 0) (filtering.ipp:13-76)."""
    data = _c(data, np.float32)
    F = data.size
    nc = degree + 1
    A = np.zeros(F * nc, np.float32)
    lib().orc_transpoly_build(C.c_uint(F), C.c_uint(degree), _p(A))
    fit = np.zeros(F, np.float32)
    coef = np.zeros(nc, np.float32)
    scratch = np.zeros(F * (nc + 1), np.float32)
    lib().orc_transpoly_eval_fit(_p(A), C.c_uint(F), C.c_uint(nc), _p(data), _p(fit), _p(coef), _p(scratch))
    return fit, coef


def get_gain(cal, T, Pss) -> float:
    cal = _c(cal, np.float32)
    return float(lib().orc_get_gain(_p(cal), C.c_float(T), C.c_float(Pss)))


def phase2(itrans, avg_final, coverage_, steady, model_temp, cal, qbar, ps, degree=6,
           exact_fit=False, ptrans_init=None):
    """Phase-2 node loop (cpp/exec/psp_process.cpp:2460-2498) + finals (:2540-2547) for a
    node slice.  Returns (pressure_transpose [n,F] f32, rms f32, avg f32, gain f32)."""
    itrans = _c(itrans, np.float32)
    n, F = itrans.shape
    ptrans = np.zeros((n, F), np.float32) if ptrans_init is None else _c(ptrans_init, np.float32).copy()
    rms = np.zeros(n, np.float64)
    avg = np.zeros(n, np.float64)
    gain = np.zeros(n, np.float64)
    cal = _c(cal, np.float32)
    args = [_c(x, np.float32) for x in (avg_final, coverage_, steady, model_temp)]
    lib().orc_phase2(n, C.c_uint(F), _p(itrans), _p(args[0]), _p(args[1]), _p(args[2]), _p(args[3]),
                     _p(cal), C.c_float(qbar), C.c_float(ps), C.c_uint(degree), int(exact_fit),
                     _p(ptrans), _p(rms), _p(avg), _p(gain))
    rms_f = np.zeros(n, np.float32)
    avg_f = np.zeros(n, np.float32)
    gain_f = np.zeros(n, np.float32)
    lib().orc_phase2_finals(_p(rms), _p(avg), _p(gain), n, C.c_uint(F), _p(rms_f), _p(avg_f), _p(gain_f))
    return ptrans, rms_f, avg_f, gain_f


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """Ask OpenMP for n threads whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


# ---------------------------------------------------------------- phase 0: projection matrix
class Camera(C.Structure):
    """orc_camera (upsp_oracle_setup.c): CameraCal's rvec/tvec/cameraMatrix/distCoeffs + frame size"""
    _fields_ = [("rvec", C.c_double * 3), ("tvec", C.c_double * 3), ("fx", C.c_double), ("fy", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double), ("k", C.c_double * 8), ("width", C.c_int), ("height", C.c_int)]


def make_camera(rvec, tvec, K, dist, width, height) -> Camera:
    cam = Camera()
    cam.rvec[:] = [float(v) for v in rvec]
    cam.tvec[:] = [float(v) for v in tvec]
    K = np.asarray(K, np.float64)
    cam.fx, cam.fy, cam.cx, cam.cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    d = list(np.asarray(dist, np.float64).ravel()) + [0.0] * 8
    cam.k[:] = d[:8]
    cam.width, cam.height = int(width), int(height)
    return cam


def project_points(cam: Camera, xyz: np.ndarray) -> np.ndarray:
    """cv::projectPoints as CameraCal::map_points_to_image calls it (cpp/lib/CameraCal.ipp:227-230)"""
    p = _c(xyz, np.float32).reshape(-1, 3)
    uv = np.empty((p.shape[0], 2), np.float32)
    lib().orc_project_points(C.byref(cam), _p(p), C.c_int(p.shape[0]), _p(uv))
    return uv


def cast_rays(verts, tri, rays):
    """nearest hit of rays [n,6] (origin, direction) over triangles tri [T,3] of verts [V,3]: (hit, t, prim)"""
    verts, tri, rays = _c(verts, np.float32), _c(tri, np.int32), _c(rays, np.float32)
    n = len(rays)
    hit, t, prim = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.int32)
    lib().orc_cast_rays(_p(verts), _p(tri), len(tri), _p(rays), n, _p(hit), _p(t), _p(prim))
    return hit, t, prim


def cam_center(cam: Camera) -> np.ndarray:
    c = np.empty(3, np.float32)
    lib().orc_cam_center(C.byref(cam), _p(c))
    return c


def create_projection(cam: Camera, xyz, normals, is_data, tri, oblique_thresh):
    """create_projection_mat (cpp/exec/psp_process.cpp:168-355): (code[N] pixel index or -1, uv[N,2])"""
    p = _c(xyz, np.float32).reshape(-1, 3)
    nr = _c(normals, np.float32).reshape(-1, 3)
    isd = _c(is_data, np.uint8)
    t = _c(tri, np.int32).reshape(-1, 3)
    code = np.empty(p.shape[0], np.int32)
    uv = np.empty((p.shape[0], 2), np.float32)
    lib().orc_create_projection(C.byref(cam), _p(p), _p(nr), _p(isd), C.c_int(p.shape[0]), _p(t), C.c_int(t.shape[0]),
                                C.c_float(oblique_thresh), _p(code), _p(uv))
    return code, uv


def projection_csr(code: np.ndarray):
    """the Eigen matrix setFromTriplets builds from the accepted nodes: one entry (pixel, 1.0) per row"""
    has = code >= 0
    rowptr = np.concatenate([[0], np.cumsum(has)]).astype(np.int32)
    return rowptr, code[has].astype(np.int32), np.ones(int(has.sum()), np.float32)
