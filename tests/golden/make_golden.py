"""Generates the golden vectors under tests/golden/ (run in the BUILD container only: it
needs /root/reference for the reference's own Python readers and fixtures, and cv2 -- the
same OpenCV entry points the reference's C++ calls).

    python tests/golden/make_golden.py

Outputs (small, committed):
  warp_golden.npz   cv2.warpAffine(u16, M, INTER_LINEAR|NEAREST + WARP_INVERSE_MAP) on random
                    12-bit frames x random affines (the call at cpp/lib/registration.cpp:69-72)
  warp_f32_golden.npz  same for f32 images (the warps inside cv::findTransformECC)
  mraw_golden.npz   first 49152 packed bytes of cpp/test/mraw/12bitMRAW.mraw (frame 1) and the
                    pixels the reference's Python reader (python/upsp/video/util.py
                    unpack_12bpp) decodes from them + CRC32 of both full decoded frames
  ecc_golden.npz    cv2.findTransformECC warp matrices / rho for synthetic frame pairs
"""
import os
import sys
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "python"))


def warp_golden():
    rng = np.random.default_rng(1234)
    nf, h, w = 9, 48, 64
    src = rng.integers(0, 4096, (nf, h, w)).astype(np.uint16)
    m = np.zeros((nf, 2, 3), np.float32)
    m[:, 0, 0] = m[:, 1, 1] = 1
    scale = np.array([5e-4, 5e-2, 0.3])[np.arange(nf) % 3]
    m[:, :, :2] += (rng.normal(0, 1, (nf, 2, 2)) * scale[:, None, None]).astype(np.float32)
    m[:, :, 2] = (rng.normal(0, 1, (nf, 2)) * np.array([1, 5, 40])[np.arange(nf) % 3][:, None]).astype(np.float32)
    lin = np.stack([cv2.warpAffine(src[f], m[f], (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP)
                    for f in range(nf)])
    nea = np.stack([cv2.warpAffine(src[f], m[f], (w, h), flags=cv2.INTER_NEAREST | cv2.WARP_INVERSE_MAP)
                    for f in range(nf)])
    np.savez_compressed(os.path.join(HERE, "warp_golden.npz"), src=src, m6=m.reshape(nf, 6), linear=lin,
                        nearest=nea, cv2_version=cv2.__version__)
    src32 = (src.astype(np.float32) + rng.random((nf, h, w)).astype(np.float32))
    lin32 = np.stack([cv2.warpAffine(src32[f], m[f], (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP)
                      for f in range(nf)])
    np.savez_compressed(os.path.join(HERE, "warp_f32_golden.npz"), src=src32, m6=m.reshape(nf, 6), linear=lin32)


def mraw_golden():
    from upsp.video import util
    raw = np.fromfile(os.path.join(REF, "cpp/test/mraw/12bitMRAW.mraw"), dtype=np.uint8)
    npix = 1024 * 1024
    fb = npix * 3 // 2
    dec = [np.asarray(util.unpack_12bpp(raw[i * fb:(i + 1) * fb].tobytes())) for i in range(2)]
    head = raw[:49152].copy()
    np.savez_compressed(os.path.join(HERE, "mraw_golden.npz"), packed_head=head,
                        pixels_head=dec[0][:32768].astype(np.uint16),
                        crc_frames=np.array([zlib.crc32(d.astype(np.uint16).tobytes()) for d in dec], np.uint64),
                        crc_packed=np.array([zlib.crc32(raw[i * fb:(i + 1) * fb].tobytes()) for i in range(2)], np.uint64))


def ecc_golden():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import upsp_b200
    h, w = 96, 128
    frames, shifts = upsp_b200.synth.make_frames(6, h, w, seed=77, hot_frames=0.0)
    ref32 = frames[0].astype(np.float32)
    ms, rhos = [], []
    for f in range(1, 6):
        M = np.eye(2, 3, dtype=np.float32)
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 50, 1e-3)
        rho, M = cv2.findTransformECC(ref32, frames[f].astype(np.float32), M, cv2.MOTION_AFFINE, crit, None, 5)
        ms.append(M.reshape(6))
        rhos.append(rho)
    np.savez_compressed(os.path.join(HERE, "ecc_golden.npz"), frames=frames, m6=np.array(ms, np.float32),
                        rho=np.array(rhos, np.float64), shifts=shifts)


if __name__ == "__main__":
    warp_golden()
    mraw_golden()
    ecc_golden()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
