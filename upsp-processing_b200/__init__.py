"""upsp-processing_b200 -- B200-native psp_process frame chain.

Python host side over the C ABI of ``libupsp_gpu.so`` (``include/upsp_gpu.h``).  The
directory name carries a hyphen (it mirrors the reference's project name), so import it
through ``upsp_b200.py`` at the repository root or ``importlib`` -- see ``load()`` there.

There is NO CPU fallback: every compute call goes to the CUDA library, and loading fails
loudly when the library is missing or no GPU is visible.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libupsp_gpu.so")

PIX_U16, PIX_PACKED12, PIX_PACKED10 = 0, 1, 2
REG_NONE, REG_PIXEL, REG_GIVEN = 0, 1, 2
INTERP_NEAREST, INTERP_LINEAR = 0, 1
PATCH_NONE, PATCH_POLYNOMIAL = 0, 1
FILTER_NONE, FILTER_GAUSSIAN, FILTER_BOX = 0, 1, 2
XCHG_PEER, XCHG_NCCL = 0, 1
NCCL_ID_BYTES = 128
IPC_HANDLE_BYTES = 64


class UpspGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"upsp_gpu error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [("device", C.c_int), ("n_cams", C.c_int), ("n_nodes", C.c_int),
                ("n_frames_total", C.c_int), ("rank", C.c_int), ("n_ranks", C.c_int),
                ("frame_capacity", C.c_int), ("batch_frames", C.c_int),
                ("pressure_aliases_intensity", C.c_int), ("keep_frame_major", C.c_int)]


class _Phase2Params(C.Structure):
    _fields_ = [("paint_cal", C.c_float * 6), ("qbar", C.c_float), ("ps", C.c_float),
                ("degree", C.c_int)]


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python upsp-processing_b200/build.py` "
                "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.upsp_gpu_last_error.restype = C.c_char_p
    return _lib


def _chk(rc):
    if rc != 0:
        raise UpspGpuError(rc, lib().upsp_gpu_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def device_count() -> int:
    return int(lib().upsp_gpu_device_count())


class PspGpu:
    """One GPU's (one rank's) share of phase 1 + transpose + phase 2.

    Mirrors the call order of the reference's phase1()/phase2()
    (cpp/exec/psp_process.cpp:1438-2043, :2262-2622)."""

    def __init__(self, n_cams, n_nodes, n_frames_total, *, device=0, rank=0, n_ranks=1,
                 frame_capacity=0, batch_frames=0, pressure_aliases_intensity=True,
                 keep_frame_major=False):
        cfg = _Config(device, n_cams, n_nodes, n_frames_total, rank, n_ranks, frame_capacity,
                      batch_frames, int(pressure_aliases_intensity), int(keep_frame_major))
        self._h = C.c_void_p()
        _chk(lib().upsp_gpu_create(C.byref(cfg), C.byref(self._h)))
        self.n_cams, self.n_nodes, self.n_frames_total = n_cams, n_nodes, n_frames_total
        s = [C.c_int() for _ in range(4)]
        _chk(lib().upsp_gpu_get_slices(self._h, *[C.byref(x) for x in s]))
        self.first_frame, self.n_frames, self.first_node, self.n_local_nodes = (x.value for x in s)

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().upsp_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- setup
    def set_camera(self, cam, width, height):
        _chk(lib().upsp_gpu_set_camera(self._h, cam, width, height))

    def set_projection(self, cam, rowptr, col, val):
        rowptr, col, val = _c(rowptr, np.int32), _c(col, np.int32), _c(val, np.float32)
        if rowptr.size != self.n_nodes + 1:
            raise ValueError("rowptr must have n_nodes+1 entries")
        _chk(lib().upsp_gpu_set_projection(self._h, cam, _p(rowptr), _p(col), _p(val)))

    def set_overlap_remap(self, src_index):
        a = _c(src_index, np.int32) if src_index is not None else None
        _chk(lib().upsp_gpu_set_overlap_remap(self._h, _p(a)))

    def set_options(self, registration=REG_NONE, interp=INTERP_LINEAR, patcher=PATCH_NONE,
                    hot_pixel_fix=True):
        _chk(lib().upsp_gpu_set_options(self._h, registration, interp, patcher, int(hot_pixel_fix)))

    def set_filter(self, kind, ksize):
        """kind: 0 none, 1 gaussian, 2 box (deck @options filter / filter_size)."""
        _chk(lib().upsp_gpu_set_filter(self._h, int(kind), int(ksize)))

    def set_patches(self, cam, bounds_off, bx, by, internal_off, ix, iy):
        bo, io = _c(bounds_off, np.int32), _c(internal_off, np.int32)
        bx, by = _c(bx, np.uint32), _c(by, np.uint32)
        ix, iy = _c(ix, np.uint32), _c(iy, np.uint32)
        _chk(lib().upsp_gpu_set_patches(self._h, cam, bo.size - 1, _p(bo), _p(bx), _p(by), _p(io),
                                        _p(ix), _p(iy)))

    def set_unpack_lut(self, lut):
        a = _c(lut, np.uint16) if lut is not None else None
        _chk(lib().upsp_gpu_set_unpack_lut(self._h, _p(a)))

    def set_reference_frame(self, cam, frame_u16):
        _chk(lib().upsp_gpu_set_reference_frame(self._h, cam, _p(_c(frame_u16, np.uint16))))

    def set_warp_matrices(self, cam, local_offset, m6):
        m = _c(m6, np.float32).reshape(-1, 6)
        _chk(lib().upsp_gpu_set_warp_matrices(self._h, cam, local_offset, m.shape[0], _p(m)))

    # -- phase 1
    def push_frames(self, cam, frames, fmt=PIX_U16, local_offset=0, count=None):
        """frames: u16 [count, H, W] (PIX_U16) or u8 [count, frame_bytes] packed; may be a raw
        integer address of pinned host memory when `count` is given."""
        if isinstance(frames, int):
            _chk(lib().upsp_gpu_push_frames(self._h, cam, C.c_void_p(frames), fmt, local_offset, count))
            return
        a = _c(frames, np.uint16 if fmt == PIX_U16 else np.uint8)
        n = a.shape[0] if count is None else count
        _chk(lib().upsp_gpu_push_frames(self._h, cam, _p(a), fmt, local_offset, n))
        self.sync()  # the numpy buffer may be pageable / temporary

    def process_frames(self, local_offset=0, count=None):
        n = self.n_frames - local_offset if count is None else count
        _chk(lib().upsp_gpu_process_frames(self._h, local_offset, n))

    def finish_phase1(self):
        _chk(lib().upsp_gpu_finish_phase1(self._h))

    def transpose(self):
        _chk(lib().upsp_gpu_transpose(self._h))

    # -- phase 2
    def phase2(self, paint_cal, qbar, ps, steady, model_temp, degree=6):
        prm = _Phase2Params()
        for i, v in enumerate(paint_cal):
            prm.paint_cal[i] = float(v)
        prm.qbar, prm.ps, prm.degree = float(qbar), float(ps), int(degree)
        st, mt = _c(steady, np.float32), _c(model_temp, np.float32)
        if st.size != self.n_nodes or mt.size != self.n_nodes:
            raise ValueError("steady / model_temp must have n_nodes entries")
        _chk(lib().upsp_gpu_phase2(self._h, C.byref(prm), _p(st), _p(mt)))

    # -- results
    def sync(self):
        _chk(lib().upsp_gpu_sync(self._h))

    def read_intensity(self, local_off=0, n=None):
        n = self.n_frames - local_off if n is None else n
        out = np.empty((n, self.n_nodes), np.float32)
        _chk(lib().upsp_gpu_read_intensity(self._h, local_off, n, _p(out)))
        return out

    def read_intensity_transpose(self, local_off=0, n=None, out=None):
        n = self.n_local_nodes - local_off if n is None else n
        out = np.empty((n, self.n_frames_total), np.float32) if out is None else out
        _chk(lib().upsp_gpu_read_intensity_transpose(self._h, local_off, n, _p(out)))
        return out

    def read_pressure_transpose(self, local_off=0, n=None, out=None):
        n = self.n_local_nodes - local_off if n is None else n
        out = np.empty((n, self.n_frames_total), np.float32) if out is None else out
        _chk(lib().upsp_gpu_read_pressure_transpose(self._h, local_off, n, _p(out)))
        return out

    def read_raw(self, which, local_off, n, host_ptr):
        """D2H into a caller-owned (e.g. pinned) buffer given by address."""
        fn = {"intensity_transpose": lib().upsp_gpu_read_intensity_transpose,
              "pressure_transpose": lib().upsp_gpu_read_pressure_transpose}[which]
        _chk(fn(self._h, local_off, n, C.c_void_p(host_ptr)))

    def read_intensity_transpose_block_async(self, local_node_off, n_nodes, frame_off, n_frames,
                                             host_ptr, host_pitch):
        _chk(lib().upsp_gpu_read_intensity_transpose_block_async(
            self._h, local_node_off, n_nodes, frame_off, n_frames, C.c_void_p(host_ptr),
            C.c_size_t(host_pitch)))

    def wait_reads(self):
        _chk(lib().upsp_gpu_wait_reads(self._h))

    def wait_pushes(self):
        """Every push_frames issued so far has read its host buffer (processing may still be running)."""
        _chk(lib().upsp_gpu_wait_pushes(self._h))

    def read_phase1_stats(self):
        avg, rms, cov = (np.empty(self.n_nodes, np.float32) for _ in range(3))
        _chk(lib().upsp_gpu_read_phase1_stats(self._h, _p(avg), _p(rms), _p(cov)))
        return avg, rms, cov

    def read_phase2_stats(self):
        rms, avg, gain = (np.empty(self.n_local_nodes, np.float32) for _ in range(3))
        _chk(lib().upsp_gpu_read_phase2_stats(self._h, _p(rms), _p(avg), _p(gain)))
        return rms, avg, gain

    def read_warp_matrices(self, cam, local_off=0, n=None, with_ecc=False):
        n = self.n_frames - local_off if n is None else n
        m = np.empty((n, 6), np.float32)
        rho = np.empty(n, np.float32) if with_ecc else None
        it = np.empty(n, np.int32) if with_ecc else None
        _chk(lib().upsp_gpu_read_warp_matrices(self._h, cam, local_off, n, _p(m), _p(rho), _p(it)))
        return (m, rho, it) if with_ecc else m

    def stage_ms(self, stage) -> float:
        ms = C.c_float()
        _chk(lib().upsp_gpu_stage_ms(self._h, stage, C.byref(ms)))
        return ms.value

    def timer_start(self):
        _chk(lib().upsp_gpu_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _chk(lib().upsp_gpu_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_kernel_sampling(self, every):
        _chk(lib().upsp_gpu_set_kernel_sampling(self._h, int(every)))

    def kernel_ms(self, kernel_class):
        ms, n = C.c_float(), C.c_int()
        _chk(lib().upsp_gpu_kernel_ms(self._h, kernel_class, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def reset_run(self):
        _chk(lib().upsp_gpu_reset_run(self._h))

    def projection_mode(self) -> int:
        m = C.c_int(-1)
        _chk(lib().upsp_gpu_projection_mode(self._h, C.byref(m)))
        return int(m.value)

    def row_bytes(self) -> int:
        b = C.c_int(4)
        _chk(lib().upsp_gpu_row_bytes(self._h, C.byref(b)))
        return int(b.value)

    def timeline(self, on=True):
        _chk(lib().upsp_gpu_timeline(self._h, int(bool(on))))

    def timeline_read(self, max_records=4096):
        rec = np.zeros((max_records, 3), np.float32)
        n = C.c_int(0)
        _chk(lib().upsp_gpu_timeline_read(self._h, _p(rec), int(max_records), C.byref(n)))
        return rec[:n.value].copy()

    def launch_count(self) -> int:
        n = C.c_longlong()
        _chk(lib().upsp_gpu_launch_count(self._h, C.byref(n)))
        return n.value

    # -- multi-GPU wiring
    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        _chk(lib().upsp_gpu_ipc_export(self._h, buf))
        return buf.raw

    def set_exchange(self, exchange):
        _chk(lib().upsp_gpu_set_exchange(self._h, int(exchange)))

    def nccl_init(self, uid: bytes):
        buf = C.create_string_buffer(bytes(uid), NCCL_ID_BYTES)
        _chk(lib().upsp_gpu_nccl_init(self._h, buf))

    def ipc_import(self, handles: bytes):
        _chk(lib().upsp_gpu_ipc_import(self._h, C.c_char_p(handles)))


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    _chk(lib().upsp_gpu_nccl_unique_id(buf))
    return buf.raw


def connect_local(ctxs):
    """Wire contexts living on different devices of THIS process (tests)."""
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    _chk(lib().upsp_gpu_connect_local(arr, len(ctxs)))


# ---------------------------------------------------------------- stand-alone operators
def op_unpack(packed, fmt, n_pixels, lut=None, device=0):
    packed = _c(packed, np.uint8)
    out = np.empty(n_pixels, np.uint16)
    l = _c(lut, np.uint16) if lut is not None else None
    _chk(lib().upsp_op_unpack(device, _p(packed), fmt, C.c_size_t(n_pixels), _p(l), _p(out)))
    return out


def op_fix_hot_pixels(frames, device=0):
    a = _c(frames, np.uint16).copy()
    nf, rows, cols = a.shape
    n_hot = np.zeros(nf, np.int32)
    _chk(lib().upsp_op_fix_hot_pixels(device, _p(a), nf, rows, cols, _p(n_hot)))
    return a, n_hot


def op_warp_affine(frames, m6, interp=INTERP_LINEAR, device=0):
    a = _c(frames, np.uint16)
    nf, h, w = a.shape
    m = _c(m6, np.float32).reshape(nf, 6)
    out = np.empty_like(a)
    _chk(lib().upsp_op_warp_affine(device, _p(a), nf, w, h, _p(m), interp, _p(out)))
    return out


def op_project_frames(rowptr, col, val, frames32, device=0):
    rowptr, col, val = _c(rowptr, np.int32), _c(col, np.int32), _c(val, np.float32)
    fr = _c(frames32, np.float32)
    nf = fr.shape[0]
    npix = fr[0].size
    n = rowptr.size - 1
    out = np.empty((nf, n), np.float32)
    _chk(lib().upsp_op_project_frames(device, _p(rowptr), _p(col), _p(val), n, _p(fr), nf,
                                      C.c_size_t(npix), _p(out)))
    return out


def op_transpose(src, device=0):
    a = _c(src, np.float32)
    y, x = a.shape
    out = np.empty((x, y), np.float32)
    _chk(lib().upsp_op_transpose(device, _p(a), x, y, _p(out)))
    return out


def op_polyfit_detrend(data, degree=6, device=0):
    a = _c(data, np.float32)
    n_pts, F = a.shape
    out = np.empty_like(a)
    _chk(lib().upsp_op_polyfit_detrend(device, _p(a), n_pts, F, degree, _p(out)))
    return out


class CameraModel(C.Structure):
    """upsp_camera_model (include/upsp_gpu.h)"""
    _fields_ = [("rvec", C.c_double * 3), ("tvec", C.c_double * 3), ("fx", C.c_double), ("fy", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double), ("dist", C.c_double * 8), ("width", C.c_int), ("height", C.c_int)]


def camera_model(rvec, tvec, K, dist, width, height) -> CameraModel:
    cam = CameraModel()
    cam.rvec[:] = [float(v) for v in rvec]
    cam.tvec[:] = [float(v) for v in tvec]
    K = np.asarray(K, np.float64)
    cam.fx, cam.fy, cam.cx, cam.cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    d = list(np.asarray(dist, np.float64).ravel()) + [0.0] * 8
    cam.dist[:] = d[:8]
    cam.width, cam.height = int(width), int(height)
    return cam


def op_project_points(cam: CameraModel, xyz, device=0):
    p = _c(xyz, np.float32).reshape(-1, 3)
    uv = np.empty((p.shape[0], 2), np.float32)
    _chk(lib().upsp_op_project_points(device, C.byref(cam), _p(p), p.shape[0], _p(uv)))
    return uv


def op_create_projection(cam: CameraModel, xyz, normals, is_datanode, tri_nodes, oblique_thresh, device=0):
    """create_projection_mat on the GPU: (code[N] = pixel index of the node's single entry or -1, uv[N,2])"""
    p = _c(xyz, np.float32).reshape(-1, 3)
    nr = _c(normals, np.float32).reshape(-1, 3)
    isd = _c(is_datanode, np.uint8)
    t = _c(tri_nodes, np.int32).reshape(-1, 3)
    code = np.empty(p.shape[0], np.int32)
    uv = np.empty((p.shape[0], 2), np.float32)
    _chk(lib().upsp_op_create_projection(device, C.byref(cam), _p(p), _p(nr), _p(isd), p.shape[0], _p(t), t.shape[0],
                                         C.c_float(oblique_thresh), _p(code), _p(uv)))
    return code, uv
