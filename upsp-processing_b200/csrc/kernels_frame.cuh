// kernels_frame.cuh -- per-frame image stages: decode (K0), hot-pixel fix (K1),
// affine warp (K3a), fiducial patch (K3b).
#pragma once
#include "common.cuh"
#include "pixel_ops.cuh"

namespace upsp {

// ---------------------------------------------------------------------------------------
// K0 + K1a: decode one batch of frames into the u16 working buffer and, in the same pass,
// find the hot pixels (>= thresh) of every frame.
// Reference: unpack_12bit / unpack_10bit cpp/lib/PSPVideo.cpp:111-150; the scan half of
// fix_hot_pixels cpp/utils/cv_extras.cpp:238-248.
// Layout: in = [frames][frame_bytes] packed, out = [frames][npix] u16.  One thread decodes
// 8 pixels (12 packed bytes -> one 16-byte store).  grid = (ceil(npix/8/256), frames).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void fix_hot_frame(uint16_t* __restrict__ img, int rows, int cols,
                                              int n, int* __restrict__ pos, int min_change) {
  int loc[UPSP_HOT_STORE];
  for (int i = 0; i < n; ++i) loc[i] = pos[i];
  for (int i = 1; i < n; ++i) {  // insertion sort -> raster order
    int v = loc[i], j = i - 1;
    while (j >= 0 && loc[j] > v) {
      loc[j + 1] = loc[j];
      --j;
    }
    loc[j + 1] = v;
  }
  for (int h = 0; h < n; ++h) {
    int row = loc[h] / cols, col = loc[h] % cols;
    int vals[4], nv = 0;
    if (row > 0) vals[nv++] = img[(size_t)(row - 1) * cols + col];
    if (col > 0) vals[nv++] = img[(size_t)row * cols + col - 1];
    if (row < rows - 1) vals[nv++] = img[(size_t)(row + 1) * cols + col];
    if (col < cols - 1) vals[nv++] = img[(size_t)row * cols + col + 1];
    for (int i = 1; i < nv; ++i) {
      int v = vals[i], j = i - 1;
      while (j >= 0 && vals[j] > v) {
        vals[j + 1] = vals[j];
        --j;
      }
      vals[j + 1] = v;
    }
    int old_val = img[(size_t)row * cols + col];
    int new_val = vals[nv / 2];
    if (old_val - new_val > min_change) img[(size_t)row * cols + col] = (uint16_t)new_val;
    pos[h] = loc[h];
  }
}

// Folded K1b: the block that finishes a frame last (ticket counter) applies that frame's
// (<= 5) hot-pixel fixes, so no separate launch is needed.  done == nullptr disables it.
__device__ __forceinline__ void hot_fix_tail(uint16_t* __restrict__ frame, int rows, int cols,
                                             int* __restrict__ hot_cnt, int* __restrict__ hot_pos,
                                             int* __restrict__ done, int f, int max_hot) {
  if (done == nullptr) return;
  __syncthreads();          // every thread's stores precede thread 0's fence (CTA causality)
  if (threadIdx.x != 0) return;
  __threadfence();          // cumulative: the block's pixels + hot notes are visible device-wide
  if (atomicAdd(done + f, 1) == (int)gridDim.x - 1) {
    __threadfence();
    const int n = *((volatile int*)(hot_cnt + f));
    if (n > 0 && n <= max_hot)
      fix_hot_frame(frame, rows, cols, n, hot_pos + f * UPSP_HOT_STORE, UPSP_HOT_MIN_CHANGE);
  }
}

__global__ void __launch_bounds__(256)
k_unpack12_scan(const uint8_t* __restrict__ in, size_t in_stride, uint16_t* __restrict__ out,
                size_t npix, int thresh, int* __restrict__ hot_cnt, int* __restrict__ hot_pos,
                int* __restrict__ done, int rows, int cols) {
  const int f = blockIdx.y;
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 8 px
  const size_t p0 = g * 8;
  const uint8_t* src = in + (size_t)f * in_stride;
  uint16_t* dst = out + (size_t)f * npix;
  if (p0 >= npix) {
  } else if (p0 + 8 <= npix && ((in_stride | (size_t)(uintptr_t)in) & 3) == 0) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src + g * 12);
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    // every 3 bytes b0 b1 b2 hold px_even = (b0<<8|b1)>>4 and px_odd = (b1<<8|b2)&0xFFF: one byte
    // permute builds (b0<<8|b1) | (b1<<8|b2)<<16, one shift + two masks finish the pair.
    const uint32_t v0 = __byte_perm(w0, w1, 0x1201);   // bytes 0,1,2
    const uint32_t v1 = __byte_perm(w0, w1, 0x4534);   // bytes 3,4,5
    const uint32_t v2 = __byte_perm(w1, w2, 0x3423);   // bytes 6,7,8  (w1 bytes 2,3 ; w2 byte 0)
    const uint32_t v3 = __byte_perm(w2, w2, 0x2312);   // bytes 9,10,11
    uint4 o;
    o.x = ((v0 >> 4) & 0x00000FFFu) | (v0 & 0x0FFF0000u);
    o.y = ((v1 >> 4) & 0x00000FFFu) | (v1 & 0x0FFF0000u);
    o.z = ((v2 >> 4) & 0x00000FFFu) | (v2 & 0x0FFF0000u);
    o.w = ((v3 >> 4) & 0x00000FFFu) | (v3 & 0x0FFF0000u);
    *reinterpret_cast<uint4*>(dst + p0) = o;
    // packed >= test (values < 2^15, thresh <= 2^15): bit 15 of (px | 0x8000) - thresh
    const uint32_t t2 = (uint32_t)min(thresh, 0x8000) * 0x00010001u;
    const uint32_t hot = (((o.x | 0x80008000u) - t2) | ((o.y | 0x80008000u) - t2) |
                          ((o.z | 0x80008000u) - t2) | ((o.w | 0x80008000u) - t2)) & 0x80008000u;
    if (hot) {
      const uint32_t px[8] = {o.x & 0xFFFF, o.x >> 16, o.y & 0xFFFF, o.y >> 16,
                              o.z & 0xFFFF, o.z >> 16, o.w & 0xFFFF, o.w >> 16};
#pragma unroll
      for (int k = 0; k < 8; ++k)
        note_hot(px[k], p0 + k, thresh, hot_cnt + f, hot_pos + f * UPSP_HOT_STORE);
    }
  } else {
    for (size_t p = p0; p < npix && p < p0 + 8; p += 2) {
      size_t b = (p >> 1) * 3;
      uint32_t x = src[b], y = src[b + 1], z = src[b + 2];
      uint32_t a0 = (x << 4) | (y >> 4), a1 = ((y & 0xF) << 8) | z;
      dst[p] = (uint16_t)a0;
      note_hot(a0, p, thresh, hot_cnt + f, hot_pos + f * UPSP_HOT_STORE);
      if (p + 1 < npix) {
        dst[p + 1] = (uint16_t)a1;
        note_hot(a1, p + 1, thresh, hot_cnt + f, hot_pos + f * UPSP_HOT_STORE);
      }
    }
  }
  hot_fix_tail(dst, rows, cols, hot_cnt, hot_pos, done, f, UPSP_HOT_MAX);
}

__device__ __noinline__ void fix_hot_frame_noinline(uint16_t* img, int rows, int cols, int n, int* pos, int min_change) {
  fix_hot_frame(img, rows, cols, n, pos, min_change);
}

__device__ __noinline__ void note_hot_item(const uint16_t* px, size_t pix0, int thresh, int* cnt, int* pos) {
  for (int k = 0; k < 32; ++k) note_hot(px[k], pix0 + k, thresh, cnt, pos);   // the thread's own stores
}

// Persistent variant of k_unpack12_scan for the two-stream pipeline: the decode of batch i+1 runs
// on a second (high-priority) stream WHILE the fused projection of batch i -- which is bound by
// instruction issue, not by HBM -- occupies most of every SM.  A small grid of long-lived blocks
// (2 per SM) streams the packed frames with 128-bit loads: one thread-iteration = 48 packed bytes
// -> 32 pixels -> four 16-byte stores, the next iteration's loads already in flight.  Block b
// owns the contiguous item range [b*chunk, (b+1)*chunk) of the batch (item = 32 pixels).
// Hot-pixel hand-over: `done[f]` counts finished items of frame f; the block that completes a
// frame applies its fixes.  Requires npix % 32 == 0, 16-byte aligned frames (host-checked).
__global__ void __launch_bounds__(256, 6)
k_unpack12_scan_p(const uint8_t* __restrict__ in, size_t in_stride, uint16_t* __restrict__ out,
                  size_t npix, int nframes, int thresh, int* __restrict__ hot_cnt,
                  int* __restrict__ hot_pos, int* __restrict__ done, int rows, int cols) {
  const unsigned ipf = (unsigned)(npix / 32);                   // items per frame
  const unsigned total = ipf * (unsigned)nframes;               // < 2^31 (host-checked)
  const unsigned chunk = (total + gridDim.x - 1) / gridDim.x;
  unsigned it0 = blockIdx.x * chunk;
  const unsigned it1 = min(total, it0 + chunk);
  const uint32_t t2 = (uint32_t)min(thresh, 0x8000) * 0x00010001u;
  while (it0 < it1) {
    const unsigned f = it0 / ipf;
    const unsigned fbeg = f * ipf;
    const unsigned seg_end = min(it1, fbeg + ipf);
    const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)f * in_stride);
    uint4* dst4 = reinterpret_cast<uint4*>(out + (size_t)f * npix);
    const unsigned iend = seg_end - fbeg;
    for (unsigned i = it0 - fbeg + threadIdx.x; i < iend; i += 256) {
      const uint4 b0 = ld_stream_u4(src + 3 * i), b1 = ld_stream_u4(src + 3 * i + 1), b2 = ld_stream_u4(src + 3 * i + 2);
      const uint32_t w[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
      uint32_t hot = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 o;
        unpack12_x8(w[3 * k], w[3 * k + 1], w[3 * k + 2], o);
        dst4[4 * i + k] = o;
        hot |= (((o.x | 0x80008000u) - t2) | ((o.y | 0x80008000u) - t2) |
                ((o.z | 0x80008000u) - t2) | ((o.w | 0x80008000u) - t2));
      }
      if (hot & 0x80008000u)         // rare: re-read the item's 32 pixels and note the hot ones
        note_hot_item(reinterpret_cast<const uint16_t*>(dst4 + 4 * i), (size_t)i * 32, thresh, hot_cnt + f,
                      hot_pos + f * UPSP_HOT_STORE);
    }
    if (done != nullptr) {
      __syncthreads();          // every thread's stores precede thread 0's fence (CTA causality)
      if (threadIdx.x == 0) {
        __threadfence();
        const int items = (int)(seg_end - it0);
        if (atomicAdd(done + f, items) + items == (int)ipf) {
          __threadfence();
          const int n = *((volatile int*)(hot_cnt + f));
          if (n > 0 && n <= UPSP_HOT_MAX)
            fix_hot_frame_noinline(reinterpret_cast<uint16_t*>(dst4), rows, cols, n, hot_pos + f * UPSP_HOT_STORE,
                                   UPSP_HOT_MIN_CHANGE);
        }
      }
    }
    it0 = seg_end;
  }
}

// 10-bit packed (5 bytes -> 4 px) with the optional 10->12-bit table
// (cpp/lib/CineReader.cpp:409-425).  One thread = one 5-byte group.
__global__ void __launch_bounds__(256)
k_unpack10_scan(const uint8_t* __restrict__ in, size_t in_stride, uint16_t* __restrict__ out,
                size_t npix, const uint16_t* __restrict__ lut, int thresh,
                int* __restrict__ hot_cnt, int* __restrict__ hot_pos,
                int* __restrict__ done, int rows, int cols) {
  const int f = blockIdx.y;
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t p0 = g * 4;
  uint16_t* dst = out + (size_t)f * npix;
  if (p0 < npix) {
    const uint8_t* s = in + (size_t)f * in_stride + g * 5;
    uint32_t p = s[0], q = s[1], r = s[2], t = s[3], u = s[4];
    uint32_t px[4] = {(p << 2) | (q >> 6), ((q & 0x3F) << 4) | (r >> 4),
                      ((r & 0x0F) << 6) | (t >> 2), ((t & 0x03) << 8) | u};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (p0 + k < npix) {
        uint32_t v = lut ? (uint32_t)__ldg(lut + px[k]) : px[k];
        dst[p0 + k] = (uint16_t)v;
        note_hot(v, p0 + k, thresh, hot_cnt + f, hot_pos + f * UPSP_HOT_STORE);
      }
    }
  }
  hot_fix_tail(dst, rows, cols, hot_cnt, hot_pos, done, f, UPSP_HOT_MAX);
}

// u16 container: copy into the working buffer (the reference's `copyTo(img)`,
// psp_process.cpp:1773) + scan.  One thread = 8 px.
__global__ void __launch_bounds__(256)
k_copy16_scan(const uint16_t* __restrict__ in, size_t in_stride_px, uint16_t* __restrict__ out,
              size_t npix, int thresh, int* __restrict__ hot_cnt, int* __restrict__ hot_pos,
              int* __restrict__ done, int rows, int cols) {
  const int f = blockIdx.y;
  const size_t p0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const uint16_t* src = in + (size_t)f * in_stride_px;
  uint16_t* dst = out + (size_t)f * npix;
  if (p0 >= npix) {
  } else if (p0 + 8 <= npix && ((npix | in_stride_px) & 7) == 0) {
    uint4 v = ld_stream_u4(src + p0);
    *reinterpret_cast<uint4*>(dst + p0) = v;
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t mx = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) mx = max(mx, max(w[k] & 0xFFFF, w[k] >> 16));
    if ((int)mx >= thresh) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        note_hot(w[k] & 0xFFFF, p0 + 2 * k, thresh, hot_cnt + f, hot_pos + f * UPSP_HOT_STORE);
        note_hot(w[k] >> 16, p0 + 2 * k + 1, thresh, hot_cnt + f, hot_pos + f * UPSP_HOT_STORE);
      }
    }
  } else {
    for (size_t p = p0; p < npix && p < p0 + 8; ++p) {
      uint16_t v = src[p];
      dst[p] = v;
      note_hot(v, p, thresh, hot_cnt + f, hot_pos + f * UPSP_HOT_STORE);
    }
  }
  hot_fix_tail(dst, rows, cols, hot_cnt, hot_pos, done, f, UPSP_HOT_MAX);
}

// ---------------------------------------------------------------------------------------
// K1b: the fix-up half of fix_hot_pixels (cv_extras.cpp:250-271).  <= 5 hot pixels per frame,
// applied serially in raster order (a later fix sees an earlier one).  One thread per frame.
// hot_cnt is left holding the count (> max_hot means "too many, frame untouched").
// ---------------------------------------------------------------------------------------
__global__ void k_fix_hot(uint16_t* __restrict__ frames, size_t npix, int rows, int cols,
                          int nframes, const int* __restrict__ hot_cnt, int* __restrict__ hot_pos,
                          int min_change, int max_hot) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nframes) return;
  int n = hot_cnt[f];
  if (n <= 0 || n > max_hot) return;
  fix_hot_frame(frames + (size_t)f * npix, rows, cols, n, hot_pos + f * UPSP_HOT_STORE, min_change);
}

// ---------------------------------------------------------------------------------------
// K3a: cv::warpAffine(src_u16, M, size, interp | WARP_INVERSE_MAP), BORDER_CONSTANT 0
// (cpp/lib/registration.cpp:69-72).  OpenCV's model: coordinates in int32 fixed point
// (AB_BITS 10, INTER_BITS 5) built from per-column / per-row tables, float weights,
// round-half-even to u16.  k_warp_tables builds the tables exactly as WarpAffineInvoker
// does (double products, cvRound); k_warp_affine samples.
//   tab layout per frame: [(adelta,bdelta)[W] | (X0,Y0)[H]] int32 pairs
// ---------------------------------------------------------------------------------------
__global__ void k_warp_tables(const float* __restrict__ m6, int nframes, int W, int H, int interp,
                              int* __restrict__ tab) {
  int f = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* M = m6 + (size_t)f * 6;
  int* t = tab + (size_t)f * (2 * W + 2 * H);
  const int round_delta = interp == 0 ? 512 : 16;
  if (i < W) {
    double x = (double)i;
    t[2 * i] = __double2int_rn(__dmul_rn(__dmul_rn((double)M[0], x), 1024.0));
    t[2 * i + 1] = __double2int_rn(__dmul_rn(__dmul_rn((double)M[3], x), 1024.0));
  }
  if (i < H) {
    double y = (double)i;
    t[2 * W + 2 * i] = __double2int_rn(__dmul_rn(
                           __dadd_rn(__dmul_rn((double)M[1], y), (double)M[2]), 1024.0)) + round_delta;
    t[2 * W + 2 * i + 1] = __double2int_rn(__dmul_rn(
                               __dadd_rn(__dmul_rn((double)M[4], y), (double)M[5]), 1024.0)) + round_delta;
  }
}

__device__ __forceinline__ uint32_t sat_u16_rn(float v) {   // v in [0, 65536): taps are u16, weights sum to 1
  int iv = f2i_rn_small(v);
  return (uint32_t)min(max(iv, 0), 65535);
}

// One pixel of the registered frame: cv::warpAffine(...)(y,x) as an integer-valued float
// (already rounded half-to-even and saturated to u16).  Used where only a few pixels of the
// registered frame are needed (patch boundary pixels, the projected nodes' pixels), so the
// registered frame itself never has to be materialised.
__device__ __forceinline__ float warp_px_u16(const uint16_t* __restrict__ s, int W, int H,
                                             const int* __restrict__ tab, int x, int y, int interp) {
  const int2 xa = __ldg(reinterpret_cast<const int2*>(tab) + x);
  const int2 ya = __ldg(reinterpret_cast<const int2*>(tab + 2 * W) + y);
  const int X = ya.x + xa.x, Y = ya.y + xa.y;
  if (interp == 0) {
    const int sx = X >> 10, sy = Y >> 10;
    return ((unsigned)sx < (unsigned)W && (unsigned)sy < (unsigned)H)
               ? u2f_exact(__ldg(s + (size_t)sy * W + sx)) : 0.0f;
  }
  const int Xs = X >> 5, Ys = Y >> 5;
  const int sx = Xs >> 5, sy = Ys >> 5;
  float v;
  if ((unsigned)sx < (unsigned)(W - 1) && (unsigned)sy < (unsigned)(H - 1)) {
    const uint16_t* p = s + (size_t)sy * W + sx;
    const float fx = frac32_exact(Xs & 31), fy = frac32_exact(Ys & 31);
    const float gx = 1.0f - fx, gy = 1.0f - fy;
    v = __fadd_rn(__fmul_rn(u2f_exact(__ldg(p)), __fmul_rn(gy, gx)),
                  __fmul_rn(u2f_exact(__ldg(p + 1)), __fmul_rn(gy, fx)));
    v = __fadd_rn(v, __fmul_rn(u2f_exact(__ldg(p + W)), __fmul_rn(fy, gx)));
    v = __fadd_rn(v, __fmul_rn(u2f_exact(__ldg(p + W + 1)), __fmul_rn(fy, fx)));
  } else {
    v = warp_sample_linear<uint16_t>(s, W, H, X, Y);
  }
  const float r = __fadd_rn(__fadd_rn(v, 12582912.0f), -12582912.0f);   // rint, half to even
  return fminf(fmaxf(r, 0.0f), 65535.0f);
}

// the same pixel from a PACKED 12-bit frame (+ its hot-pixel fix list): the decoded frame is never materialised
__device__ __forceinline__ float warp_px_p12(const uint8_t* __restrict__ fr, const HotFix* __restrict__ h, int W, int H,
                                             const int* __restrict__ tab, int x, int y, int interp) {
  const int2 xa = __ldg(reinterpret_cast<const int2*>(tab) + x);
  const int2 ya = __ldg(reinterpret_cast<const int2*>(tab + 2 * W) + y);
  const int X = ya.x + xa.x, Y = ya.y + xa.y;
  if (interp == 0) {
    const int sx = X >> 10, sy = Y >> 10;
    return ((unsigned)sx < (unsigned)W && (unsigned)sy < (unsigned)H)
               ? u2f_exact(px_packed12_fixed(fr, (unsigned)(sy * W + sx), h)) : 0.0f;
  }
  const int Xs = X >> 5, Ys = Y >> 5;
  const int sx = Xs >> 5, sy = Ys >> 5;
  const unsigned fxi = Xs & 31, fyi = Ys & 31;
  uint32_t t[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xx = sx + (k & 1), yy = sy + (k >> 1);
    t[k] = ((unsigned)xx < (unsigned)W && (unsigned)yy < (unsigned)H) ? px_packed12_fixed(fr, (unsigned)(yy * W + xx), h) : 0u;
  }
  // 12-bit taps, weights k/1024: every product and partial sum of OpenCV's float sequence is exact, so the
  // value is S/1024 rounded half to even (the same identity k_project_fused4 uses)
  const unsigned S = (t[0] * (32u - fxi) + t[1] * fxi) * (32u - fyi) + (t[2] * (32u - fxi) + t[3] * fxi) * fyi;
  return __fadd_rn(__fmaf_rn(__uint_as_float(S + 0x4B000000u), 0.0009765625f, 12574720.0f), -12582912.0f);
}

// 8 output pixels per thread (one 16-byte store).  Near-identity maps (the registration case:
// sub-pixel jitter) put the 8 pixels' taps in one contiguous 9-pixel run of two source rows;
// that run is fetched with three 8-byte loads per row instead of 32 two-byte loads.  Any other
// map (large rotation / scale, image border) takes the per-pixel path with identical
// arithmetic.  block = 128 threads = 1024 px of one row; grid (ceil(W/1024), H, frames).
__device__ __forceinline__ void window9(const uint16_t* __restrict__ p /*8B aligned*/, int off,
                                        uint32_t (&px)[9]) {
  const uint2* q = reinterpret_cast<const uint2*>(p);
  const uint2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  uint32_t w[7] = {a.x, a.y, b.x, b.y, c.x, c.y, 0u};
  const int wo = off >> 1, hs = (off & 1) * 16;
  uint32_t u[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const uint32_t lo = wo ? w[i + 1] : w[i], hi = wo ? w[i + 2] : w[i + 1];
    u[i] = __funnelshift_r(lo, hi, hs);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) px[k] = (u[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
}

__global__ void __launch_bounds__(128)
k_warp_affine8_u16(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int W, int H,
                   const int* __restrict__ tab, int interp, int skip_frame) {
  const int f = blockIdx.z;
  const int x = (blockIdx.x * 128 + threadIdx.x) * 8;
  const int y = blockIdx.y;
  if (x >= W) return;
  const size_t P = (size_t)W * H;
  const uint16_t* s = src + (size_t)f * P;
  uint16_t* d = dst + (size_t)f * P + (size_t)y * W + x;
  const bool full = (x + 8 <= W) && ((W & 7) == 0);
  if (f == skip_frame) {  // global frame 0 is never registered (psp_process.cpp:1777)
    if (full) {
      *reinterpret_cast<uint4*>(d) = __ldg(reinterpret_cast<const uint4*>(s + (size_t)y * W + x));
    } else {
      for (int k = 0; k < 8 && x + k < W; ++k) d[k] = s[(size_t)y * W + x + k];
    }
    return;
  }
  const int* t = tab + (size_t)f * (2 * W + 2 * H);
  const int2 y0 = __ldg(reinterpret_cast<const int2*>(t + 2 * W) + y);
  const int X0 = y0.x, Y0 = y0.y;
  int X[8], Y[8];
  if (full) {
    const int4* q = reinterpret_cast<const int4*>(t + 2 * x);   // (ad,bd) pairs of 8 pixels
    const int4 a0 = __ldg(q), a1 = __ldg(q + 1), a2 = __ldg(q + 2), a3 = __ldg(q + 3);
    X[0] = X0 + a0.x; Y[0] = Y0 + a0.y; X[1] = X0 + a0.z; Y[1] = Y0 + a0.w;
    X[2] = X0 + a1.x; Y[2] = Y0 + a1.y; X[3] = X0 + a1.z; Y[3] = Y0 + a1.w;
    X[4] = X0 + a2.x; Y[4] = Y0 + a2.y; X[5] = X0 + a2.z; Y[5] = Y0 + a2.w;
    X[6] = X0 + a3.x; Y[6] = Y0 + a3.y; X[7] = X0 + a3.z; Y[7] = Y0 + a3.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int xx = min(x + k, W - 1);
      X[k] = X0 + __ldg(t + 2 * xx);
      Y[k] = Y0 + __ldg(t + 2 * xx + 1);
    }
  }
  uint32_t o[8];
  bool done = false;
  if (interp == 1 && full) {
    const int sx0 = X[0] >> 10, sy0 = Y[0] >> 10;
    bool run = sx0 >= 0 && sx0 + 9 <= W && sy0 >= 0 && sy0 + 1 < H;
#pragma unroll
    for (int k = 1; k < 8; ++k) run = run && ((X[k] >> 10) == sx0 + k) && ((Y[k] >> 10) == sy0);
    const size_t a = (size_t)sy0 * W + sx0;
    const size_t a4 = a & ~(size_t)3;
    run = run && ((W & 3) == 0) && (a4 + W + 12 <= P);
    if (run) {
      uint32_t r0[9], r1[9];
      window9(s + a4, (int)(a - a4), r0);
      window9(s + a4 + W, (int)(a - a4), r1);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int Xs = X[k] >> 5, Ys = Y[k] >> 5;
        const float fx = frac32_exact(Xs & 31), fy = frac32_exact(Ys & 31);
        const float w0 = __fmul_rn(1.0f - fy, 1.0f - fx), w1 = __fmul_rn(1.0f - fy, fx);
        const float w2 = __fmul_rn(fy, 1.0f - fx), w3 = __fmul_rn(fy, fx);
        float v = __fadd_rn(__fmul_rn(u2f_exact(r0[k]), w0), __fmul_rn(u2f_exact(r0[k + 1]), w1));
        v = __fadd_rn(v, __fmul_rn(u2f_exact(r1[k]), w2));
        v = __fadd_rn(v, __fmul_rn(u2f_exact(r1[k + 1]), w3));
        o[k] = sat_u16_rn(v);
      }
      done = true;
    }
  }
  if (!done) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      o[k] = 0;
      if (x + k >= W) continue;
      if (interp == 0) {
        const int sx = X[k] >> 10, sy = Y[k] >> 10;
        o[k] = (sx >= 0 && sx < W && sy >= 0 && sy < H) ? (uint32_t)__ldg(s + (size_t)sy * W + sx) : 0u;
      } else {
        o[k] = sat_u16_rn(warp_sample_linear<uint16_t>(s, W, H, X[k], Y[k]));
      }
    }
  }
  if (full) {
    *reinterpret_cast<uint4*>(d) = make_uint4(o[0] | (o[1] << 16), o[2] | (o[3] << 16),
                                              o[4] | (o[5] << 16), o[6] | (o[7] << 16));
  } else {
    for (int k = 0; k < 8 && x + k < W; ++k) d[k] = (uint16_t)o[k];
  }
}

// ---------------------------------------------------------------------------------------
// K3b: PatchClusters<float>::operator() (cpp/lib/patches.ipp:98-164).
// The cubic 2-D Vandermonde A depends only on cluster geometry, so its float
// ColPivHouseholderQR (patches.ipp:203) is factored ONCE on the host (host_qr.hpp) and the
// per-frame work is: z = boundary pixels; c = Q^T z (apply the stored reflectors in order);
// back-substitute R; un-permute; evaluate at the interior pixels (polyval2D :208-236).
// Every float operation is issued in the reference's order with explicit round-to-nearest
// mul/add (no FMA contraction) because the system is numerically rank-deficient in float
// and the patched values are only reproducible operation-for-operation.
// One warp per (cluster, 32 frames): lane = frame, so scratch/pixel accesses of a warp are
// the same element of 32 consecutive frames.
// ---------------------------------------------------------------------------------------
struct PatchGeom {            // device pointers, one per camera
  int n_clusters;
  const int* bounds_off;      // [ncl+1]
  const int* bsrc;            // [nb_total]  >=0: pixel index; <0: -1-slot (patched value)
  const float* qr_e;          // [10*nb_total] essential part of reflector k at [10*off + k*nb + i]; R on/above diag
  const float* qr_te;         // same layout: tau_k * e_k[i]
  const float* hcoef;         // [ncl*10]
  const int* perm;            // [ncl*10]
  const int* nzp;             // [ncl] nonzero pivots (0 => cluster inactive)
  const int* internal_off;    // [ncl+1]
  const float* ipow;          // [6*ni_total]: x, x^2, x^3, y, y^2, y^3 as float
  const int* islot_pix;       // [ni_total] pixel index of each interior pixel
  const int* order;           // clusters sorted by dependency level
};

// One warp per (cluster, frame): lanes stride over the boundary / interior pixels; a block
// holds PATCH_WARPS frames of the same cluster and stages the cluster's reflectors (E, tau*E)
// in shared memory once.  The only order-sensitive reduction, s = ((p0+p1)+p2)+..., is done by
// lane 0 over products that all lanes computed in parallel, so the sequential chain is n adds
// instead of the whole solve.  tab != nullptr: the boundary pixels are taken from the REGISTERED
// frame, computed on the fly (warp_px_u16) from the decoded frame.
constexpr int PATCH_WARPS = 4;

__global__ void __launch_bounds__(32 * PATCH_WARPS)
k_patch(PatchGeom g, const int* __restrict__ cl_list, const uint16_t* __restrict__ frames,
        size_t npix, int W, int H, const int* __restrict__ tab, int interp, int skip_frame,
        int nframes, int bstride, float* __restrict__ pv, const uint8_t* __restrict__ packed = nullptr,
        size_t frame_bytes = 0, const HotFix* __restrict__ hot = nullptr) {
  extern __shared__ float psh[];
  const int cl = cl_list[blockIdx.x];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y * PATCH_WARPS + wid;
  const int nz = g.nzp[cl];
  if (nz < 0) return;
  const int off = g.bounds_off[cl], nb = g.bounds_off[cl + 1] - off;
  float* E = psh;                 // [10*nb]
  float* TE = psh + 10 * nb;      // [10*nb]
  for (int i = threadIdx.x; i < 10 * nb; i += blockDim.x) {
    E[i] = __ldg(g.qr_e + (size_t)10 * off + i);
    TE[i] = __ldg(g.qr_te + (size_t)10 * off + i);
  }
  __syncthreads();
  if (b >= nframes) return;
  float* c = psh + 20 * nb + wid * (2 * nb + 16);   // [nb]
  float* prod = c + nb;                              // [nb]
  float* poly = c + 2 * nb;                          // [10]
  const uint16_t* img = frames + (size_t)b * npix;
  const int* t = (tab != nullptr && b != skip_frame) ? tab + (size_t)b * (2 * W + 2 * H) : nullptr;
  for (int i = lane; i < nb; i += 32) {
    const int s = __ldg(g.bsrc + off + i);
    float v;
    if (s < 0) v = pv[(size_t)(-1 - s) * bstride + b];
    else if (packed != nullptr) {      // packed 12-bit source (TMA projection mode): pixels + the frame's fix list
      const uint8_t* fr = packed + (size_t)b * frame_bytes;
      const HotFix* h = hot ? hot + b : nullptr;
      v = t ? warp_px_p12(fr, h, W, H, t, s % W, s / W, interp) : u2f_exact(px_packed12_fixed(fr, (unsigned)s, h));
    }
    else if (t) v = warp_px_u16(img, W, H, t, s % W, s / W, interp);
    else v = u2f_exact(__ldg(img + s));
    c[i] = v;
  }
  __syncwarp();
  const float* hc = g.hcoef + cl * 10;
  for (int k = 0; k < nz; ++k) {
    const int n = nb - k;
    const float tau = hc[k];
    if (n == 1) {
      if (lane == 0) c[k] = __fmul_rn(c[k], __fsub_rn(1.0f, tau));
    } else if (tau != 0.0f) {
      const float* e = E + (size_t)k * nb + k + 1;
      const float* te = TE + (size_t)k * nb + k + 1;
      for (int i = lane; i < n - 1; i += 32) prod[i] = __fmul_rn(e[i], c[k + 1 + i]);
      __syncwarp();
      float tt = 0.0f;
      if (lane == 0) {
        float s = 0.0f;
        int i = 0;
        for (; i + 8 <= n - 1; i += 8) {
          const float p0 = prod[i], p1 = prod[i + 1], p2 = prod[i + 2], p3 = prod[i + 3];
          const float p4 = prod[i + 4], p5 = prod[i + 5], p6 = prod[i + 6], p7 = prod[i + 7];
          s = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s, p0), p1), p2), p3);
          s = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s, p4), p5), p6), p7);
        }
        for (; i < n - 1; ++i) s = __fadd_rn(s, prod[i]);
        tt = __fadd_rn(s, c[k]);
        c[k] = __fsub_rn(c[k], __fmul_rn(tau, tt));
      }
      tt = __shfl_sync(0xffffffffu, tt, 0);
      for (int i = lane; i < n - 1; i += 32) c[k + 1 + i] = __fsub_rn(c[k + 1 + i], __fmul_rn(tt, te[i]));
    }
    __syncwarp();
  }
  if (lane == 0) {
    float x[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) x[i] = i < nz ? c[i] : 0.0f;
#pragma unroll
    for (int i = 9; i >= 0; --i) {
      if (i < nz) {
        x[i] = __fdiv_rn(x[i], E[(size_t)i * nb + i]);
#pragma unroll
        for (int j = 0; j < 10; ++j)
          if (j < i) x[j] = __fsub_rn(x[j], __fmul_rn(x[i], E[(size_t)i * nb + j]));
      }
    }
    const int* pm = g.perm + cl * 10;
#pragma unroll
    for (int i = 0; i < 10; ++i) poly[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < 10; ++i) poly[pm[i]] = i < nz ? x[i] : 0.0f;
  }
  __syncwarp();
  float pl[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) pl[i] = poly[i];
  // polyval2D: z = sum_c poly[c] * y^i * x^j, order [1,x,x2,x3,y,xy,x2y,y2,xy2,y3]
  const int ioff = g.internal_off[cl], ni = g.internal_off[cl + 1] - ioff;
  for (int i = lane; i < ni; i += 32) {
    const float* pw = g.ipow + (size_t)6 * (ioff + i);
    const float xp[4] = {1.0f, __ldg(pw), __ldg(pw + 1), __ldg(pw + 2)};
    const float yp[4] = {1.0f, __ldg(pw + 3), __ldg(pw + 4), __ldg(pw + 5)};
    float z = 0.0f;
    int cnt = 0;
#pragma unroll
    for (int a = 0; a <= 3; ++a)
#pragma unroll
      for (int bb = 0; bb <= 3; ++bb)
        if (a + bb <= 3) {
          z = __fadd_rn(z, __fmul_rn(__fmul_rn(pl[cnt], yp[a]), xp[bb]));
          ++cnt;
        }
    pv[(size_t)(ioff + i) * bstride + b] = z;
  }
}

}  // namespace upsp
