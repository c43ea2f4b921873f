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
  synthetic10.cine / synthetic12.cine   tiny Vision Research cine files (3 frames 64x32, 10- and
                    12-bit packed, one tagged exposure block) assembled from the reference's own
                    ctypes header structures (python/upsp/video/cine.py)
  tiny12.mraw / tiny12.cih   tiny Photron pair (2 frames 32x16, 12-bit)
  setup_golden.npz  cv2.projectPoints / cv2.Rodrigues for random calibrations (phase-0 camera model)
  video_golden.npz  what the reference's Python readers (upsp.video.CineReader / MrawReader) decode
                    from those files, their properties, and the 10->12-bit table
  ../../upsp-processing_b200/host/cine_lut.inc   the same table as a C initialiser list
  ../../upsp-processing_b200/csrc/gauss_fixed.inc   OpenCV's 16.16 fixed-point Gaussian taps, sizes 3..31
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


def setup_golden():
    """cv2.projectPoints (the call of CameraCal::map_points_to_image, cpp/lib/CameraCal.ipp:227-230)
    and cv2.Rodrigues for random calibrations with 4, 5 and 8 distortion coefficients."""
    rng = np.random.default_rng(2024)
    cams, pts, uvs, rots = [], [], [], []
    for trial in range(12):
        rvec = rng.normal(0, 0.5, 3)
        tvec = np.array([rng.normal(0, 2), rng.normal(0, 2), 30 + rng.normal(0, 5)])
        K = np.array([[1200 + rng.normal(0, 50), 0, 512 + rng.normal(0, 5)], [0, 1190 + rng.normal(0, 50), 500 + rng.normal(0, 5)],
                      [0, 0, 1]])
        nd = [4, 5, 8][trial % 3]
        d = np.concatenate([rng.normal(0, 0.05, 2), rng.normal(0, 0.002, 2), rng.normal(0, 0.01, 1), rng.normal(0, 0.01, 3)])
        d[nd:] = 0.0
        p = rng.normal(0, 6, (64, 3)).astype(np.float32)
        uv = cv2.projectPoints(p.reshape(-1, 1, 3), rvec, tvec, K, d[:nd])[0].reshape(-1, 2).astype(np.float32)
        cams.append(np.concatenate([rvec, tvec, [K[0, 0], K[1, 1], K[0, 2], K[1, 2]], d]))
        pts.append(p)
        uvs.append(uv)
        rots.append(cv2.Rodrigues(rvec)[0].reshape(9))
    np.savez_compressed(os.path.join(HERE, "setup_golden.npz"), cams=np.array(cams), pts=np.array(pts), uv=np.array(uvs),
                        rot=np.array(rots), cv2_version=cv2.__version__)


def video_golden():
    """Synthetic containers written with the reference's header structures, decoded with the
    reference's Python readers."""
    import ctypes
    import struct
    from upsp.video import cine, mraw
    rng = np.random.default_rng(77)
    W, H, NF = 64, 32, 3
    out = {}
    for bpp in (10, 12):
        vals = rng.integers(0, 1 << bpp, (NF, H * W)).astype(np.uint32)
        frames = []
        for f in range(NF):                       # pack exactly as PSPVideo.cpp:111-150 unpacks
            v = vals[f]
            if bpp == 12:
                a, b = v[0::2], v[1::2]
                by = np.stack([a >> 4, ((a & 0xF) << 4) | (b >> 8), b & 0xFF], 1)
            else:
                a, b, c, d = v[0::4], v[1::4], v[2::4], v[3::4]
                by = np.stack([a >> 2, ((a & 3) << 6) | (b >> 4), ((b & 0xF) << 4) | (c >> 6),
                               ((c & 0x3F) << 2) | (d >> 8), d & 0xFF], 1)
            frames.append(by.astype(np.uint8).tobytes())
        cf, bm, st = cine.CINEFILEHEADER(), cine.BITMAPINFOHEADER(), cine.SETUP()
        st.Length = ctypes.sizeof(st)
        st.Mark = 0x5453
        st.FrameRate16 = 5000
        st.FrameRate = 5000
        st.ImWidth, st.ImHeight, st.RealBPP = W, H, bpp
        st.LensAperture = 2.8
        tagged = struct.pack("<IHH", 8 + 4 * NF, 0x3eb, 0) + struct.pack("<%dI" % NF, *([int(2 ** 32 * 25e-6)] * NF))
        cf.Type = 0x4943
        cf.Headersize = ctypes.sizeof(cf)
        cf.Version = 1
        cf.TotalImageCount = cf.ImageCount = NF
        cf.OffImageHeader = ctypes.sizeof(cf)
        cf.OffSetup = ctypes.sizeof(cf) + ctypes.sizeof(bm)
        cf.OffImageOffsets = cf.OffSetup + st.Length + len(tagged)
        bm.biSize = ctypes.sizeof(bm)
        bm.biWidth, bm.biHeight, bm.biPlanes, bm.biBitCount = W, H, 1, 16
        bm.biCompression = 256 if bpp == 10 else 1024
        bm.biSizeImage = len(frames[0])
        first = cf.OffImageOffsets + 8 * NF
        offs = [first + i * (8 + len(frames[0])) for i in range(NF)]
        path = os.path.join(HERE, "synthetic%d.cine" % bpp)
        with open(path, "wb") as fd:
            fd.write(bytes(cf) + bytes(bm) + bytes(st) + tagged + struct.pack("<%dq" % NF, *offs))
            for fr in frames:
                fd.write(struct.pack("<II", 8, len(fr)) + fr)
        with cine.CineReader(path) as rd:
            dec = np.stack([np.asarray(rd.read_frame(i)) for i in range(NF)]).astype(np.uint16)
            out["cine%d_frames" % bpp] = dec
            out["cine%d_props" % bpp] = np.array([rd.width, rd.height, rd.bit_depth, rd.frame_count, rd.frame_rate])
        out["cine%d_crc" % bpp] = np.array([zlib.crc32(fr) for fr in frames], np.uint64)
    lut = np.asarray(cine._LUT_10BIT).astype(np.uint16)
    out["lut10"] = lut
    with open(os.path.join(HERE, "..", "..", "upsp-processing_b200", "host", "cine_lut.inc"), "w") as fd:
        fd.write("// 10 -> 12-bit table of packed 10-bit cines; generated by tests/golden/make_golden.py from the\n"
                 "// reference's python/upsp/video/cine.py (_LUT_10BIT == CINE2_LUT, cpp/lib/CineReader.cpp:23-87)\n")
        for i in range(0, 1024, 16):
            fd.write(", ".join(str(int(x)) for x in lut[i:i + 16]) + ",\n")
    # Photron pair
    w, h, nf = 32, 16, 2
    vals = rng.integers(0, 4096, (nf, h * w)).astype(np.uint32)
    raw = b""
    for f in range(nf):
        a, b = vals[f][0::2], vals[f][1::2]
        raw += np.stack([a >> 4, ((a & 0xF) << 4) | (b >> 8), b & 0xFF], 1).astype(np.uint8).tobytes()
    open(os.path.join(HERE, "tiny12.mraw"), "wb").write(raw)
    open(os.path.join(HERE, "tiny12.cih"), "w", newline="").write(
        "#Camera Information Header\r\nDate : 2020/9/1\r\nCamera Type : FASTCAM synthetic\r\nScene Name : \r\n"
        "Record Rate(fps) : 1000\r\nShutter Speed(s) : 1/1000\r\nTotal Frame : %d\r\nImage Width : %d\r\n"
        "Image Height : %d\r\nColor Type : Mono\r\nColor Bit : 12\r\nFile Format : MRaw\r\nEffectiveBit Depth : 12\r\n"
        "#END\r\n" % (nf, w, h))
    with mraw.MrawReader(os.path.join(HERE, "tiny12.mraw")) as rd:
        out["mraw_frames"] = np.stack([np.asarray(rd.read_frame(i)) for i in range(nf)]).astype(np.uint16)
        out["mraw_props"] = np.array([rd.width, rd.height, rd.bit_depth, rd.frame_count, rd.frame_rate])
    out["mraw_crc"] = np.array([zlib.crc32(raw[i * len(raw) // nf:(i + 1) * len(raw) // nf]) for i in range(nf)], np.uint64)
    np.savez_compressed(os.path.join(HERE, "video_golden.npz"), **out)


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


def gauss_fixed_taps(k):
    """OpenCV's 16.16 fixed-point taps of GaussianBlur(CV_16U, Size(k,k), sigma=0): the bit-exact double kernel
    (cv2.getGaussianKernel) rounded with error diffusion, centre tap = 65536 - the rest (getGaussianKernelFixedPoint_ED)."""
    kern = cv2.getGaussianKernel(k, 0, cv2.CV_64F).ravel()
    half, err, total = [], 0.0, 0
    for i in range(k // 2):
        adj = kern[i] * 65536.0 + err
        v0 = int(np.rint(adj))                 # cvRound
        err = adj - v0
        half.append(v0)
        total += v0
    return half + [65536 - 2 * total]


def gauss_golden():
    lines = ["// gauss_fixed.inc -- OpenCV's 16.16 fixed-point Gaussian taps for sigma = 0 (GaussianBlur on CV_16U: getGaussianKernelBitExact +",
             "// getGaussianKernelFixedPoint_ED, error-diffusion rounding, centre tap = 65536 - the rest), first half + centre per odd size 3..31.",
             "// Generated by tests/golden/make_golden.py gauss (cv2.getGaussianKernel(k, 0, CV_64F) -> the rounding above); tests/test_oracle_golden.py",
             "// holds the integer filter built from these taps bit-exact against cv2.GaussianBlur for every size.",
             "static const int kGaussFixed[15][16] = {"]
    for k in range(3, 33, 2):
        lines.append("  {" + ", ".join(map(str, gauss_fixed_taps(k))) + "},   // k = %d" % k)
    lines.append("};")
    with open(os.path.join(HERE, "..", "..", "upsp-processing_b200", "csrc", "gauss_fixed.inc"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gauss":     # only the fixed-point Gaussian taps
        gauss_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "video":     # only the container fixtures
        video_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "setup":     # only the camera-model fixtures
        setup_golden()
        sys.exit(0)
    warp_golden()
    mraw_golden()
    ecc_golden()
    video_golden()
    setup_golden()
    gauss_golden()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
