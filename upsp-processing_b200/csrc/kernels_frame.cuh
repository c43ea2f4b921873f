// kernels_frame.cuh -- per-frame image stages: decode (K0), hot-pixel fix (K1),
// affine warp (K3a), fiducial patch (K3b).
#pragma once
#include "common.cuh"

namespace upsp {

// ---------------------------------------------------------------------------------------
// K0 + K1a: decode one batch of frames into the u16 working buffer and, in the same pass,
// find the hot pixels (>= thresh) of every frame.
// Reference: unpack_12bit / unpack_10bit cpp/lib/PSPVideo.cpp:111-150; the scan half of
// fix_hot_pixels cpp/utils/cv_extras.cpp:238-248.
// Layout: in = [frames][frame_bytes] packed, out = [frames][npix] u16.  One thread decodes
// 8 pixels (12 packed bytes -> one 16-byte store).  grid = (ceil(npix/8/256), frames).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void note_hot(uint32_t v, size_t pix, int thresh, int* cnt, int* pos) {
  if ((int)v >= thresh) {
    int s = atomicAdd(cnt, 1);
    if (s < UPSP_HOT_STORE) pos[s] = (int)pix;
  }
}

__global__ void __launch_bounds__(256)
k_unpack12_scan(const uint8_t* __restrict__ in, size_t in_stride, uint16_t* __restrict__ out,
                size_t npix, int thresh, int* __restrict__ hot_cnt, int* __restrict__ hot_pos) {
  const int f = blockIdx.y;
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 8 px
  const size_t p0 = g * 8;
  if (p0 >= npix) return;
  const uint8_t* src = in + (size_t)f * in_stride;
  uint16_t* dst = out + (size_t)f * npix;
  if (p0 + 8 <= npix && ((in_stride | (size_t)(uintptr_t)in) & 3) == 0) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src + g * 12);
    uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    // bytes b0..b11, MSB-first 12-bit fields: px0=b0<<4|b1>>4, px1=(b1&15)<<8|b2, ...
    uint32_t b0 = w0 & 0xFF, b1 = (w0 >> 8) & 0xFF, b2 = (w0 >> 16) & 0xFF, b3 = w0 >> 24;
    uint32_t b4 = w1 & 0xFF, b5 = (w1 >> 8) & 0xFF, b6 = (w1 >> 16) & 0xFF, b7 = w1 >> 24;
    uint32_t b8 = w2 & 0xFF, b9 = (w2 >> 8) & 0xFF, b10 = (w2 >> 16) & 0xFF, b11 = w2 >> 24;
    uint32_t px[8];
    px[0] = (b0 << 4) | (b1 >> 4);
    px[1] = ((b1 & 0xF) << 8) | b2;
    px[2] = (b3 << 4) | (b4 >> 4);
    px[3] = ((b4 & 0xF) << 8) | b5;
    px[4] = (b6 << 4) | (b7 >> 4);
    px[5] = ((b7 & 0xF) << 8) | b8;
    px[6] = (b9 << 4) | (b10 >> 4);
    px[7] = ((b10 & 0xF) << 8) | b11;
    uint4 o = make_uint4(px[0] | (px[1] << 16), px[2] | (px[3] << 16), px[4] | (px[5] << 16),
                         px[6] | (px[7] << 16));
    *reinterpret_cast<uint4*>(dst + p0) = o;
    uint32_t mx = max(max(max(px[0], px[1]), max(px[2], px[3])),
                      max(max(px[4], px[5]), max(px[6], px[7])));
    if ((int)mx >= thresh) {
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
}

// 10-bit packed (5 bytes -> 4 px) with the optional 10->12-bit table
// (cpp/lib/CineReader.cpp:409-425).  One thread = one 5-byte group.
__global__ void __launch_bounds__(256)
k_unpack10_scan(const uint8_t* __restrict__ in, size_t in_stride, uint16_t* __restrict__ out,
                size_t npix, const uint16_t* __restrict__ lut, int thresh,
                int* __restrict__ hot_cnt, int* __restrict__ hot_pos) {
  const int f = blockIdx.y;
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t p0 = g * 4;
  if (p0 >= npix) return;
  const uint8_t* s = in + (size_t)f * in_stride + g * 5;
  uint32_t p = s[0], q = s[1], r = s[2], t = s[3], u = s[4];
  uint32_t px[4] = {(p << 2) | (q >> 6), ((q & 0x3F) << 4) | (r >> 4),
                    ((r & 0x0F) << 6) | (t >> 2), ((t & 0x03) << 8) | u};
  uint16_t* dst = out + (size_t)f * npix;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (p0 + k < npix) {
      uint32_t v = lut ? (uint32_t)__ldg(lut + px[k]) : px[k];
      dst[p0 + k] = (uint16_t)v;
      note_hot(v, p0 + k, thresh, hot_cnt + f, hot_pos + f * UPSP_HOT_STORE);
    }
  }
}

// u16 container: copy into the working buffer (the reference's `copyTo(img)`,
// psp_process.cpp:1773) + scan.  One thread = 8 px.
__global__ void __launch_bounds__(256)
k_copy16_scan(const uint16_t* __restrict__ in, size_t in_stride_px, uint16_t* __restrict__ out,
              size_t npix, int thresh, int* __restrict__ hot_cnt, int* __restrict__ hot_pos) {
  const int f = blockIdx.y;
  const size_t p0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (p0 >= npix) return;
  const uint16_t* src = in + (size_t)f * in_stride_px;
  uint16_t* dst = out + (size_t)f * npix;
  if (p0 + 8 <= npix && ((npix | in_stride_px) & 7) == 0) {
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
  int* pos = hot_pos + f * UPSP_HOT_STORE;
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
  uint16_t* img = frames + (size_t)f * npix;
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

// ---------------------------------------------------------------------------------------
// K3a: cv::warpAffine(src_u16, M, size, interp | WARP_INVERSE_MAP), BORDER_CONSTANT 0
// (cpp/lib/registration.cpp:69-72).  OpenCV's model: coordinates in int32 fixed point
// (AB_BITS 10, INTER_BITS 5) built from per-column / per-row tables, float weights,
// round-half-even to u16.  k_warp_tables builds the tables exactly as WarpAffineInvoker
// does (double products, cvRound); k_warp_affine samples.
//   tab layout per frame: [adelta[W] | bdelta[W] | X0[H] | Y0[H]] int32
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
    t[i] = __double2int_rn(__dmul_rn(__dmul_rn((double)M[0], x), 1024.0));
    t[W + i] = __double2int_rn(__dmul_rn(__dmul_rn((double)M[3], x), 1024.0));
  }
  if (i < H) {
    double y = (double)i;
    t[2 * W + i] = __double2int_rn(__dmul_rn(
                       __dadd_rn(__dmul_rn((double)M[1], y), (double)M[2]), 1024.0)) + round_delta;
    t[2 * W + H + i] = __double2int_rn(__dmul_rn(
                           __dadd_rn(__dmul_rn((double)M[4], y), (double)M[5]), 1024.0)) + round_delta;
  }
}

template <typename LoadT>
__device__ __forceinline__ float warp_sample_linear(const LoadT* __restrict__ src, int W, int H,
                                                    int X, int Y) {
  X >>= 5;
  Y >>= 5;
  const int sx = X >> 5, sy = Y >> 5;
  const float fx = (float)(X & 31) * 0.03125f, fy = (float)(Y & 31) * 0.03125f;
  if (sx >= W || sx + 1 < 0 || sy >= H || sy + 1 < 0) return 0.0f;
  const float w0 = __fmul_rn(1.0f - fy, 1.0f - fx), w1 = __fmul_rn(1.0f - fy, fx);
  const float w2 = __fmul_rn(fy, 1.0f - fx), w3 = __fmul_rn(fy, fx);
  const bool x0 = sx >= 0, x1 = sx + 1 < W, y0 = sy >= 0, y1 = sy + 1 < H;
  const LoadT* r0 = src + (size_t)(y0 ? sy : 0) * W;
  const LoadT* r1 = src + (size_t)(y1 ? sy + 1 : 0) * W;
  float v0 = (x0 && y0) ? (float)__ldg(r0 + sx) : 0.0f;
  float v1 = (x1 && y0) ? (float)__ldg(r0 + sx + 1) : 0.0f;
  float v2 = (x0 && y1) ? (float)__ldg(r1 + sx) : 0.0f;
  float v3 = (x1 && y1) ? (float)__ldg(r1 + sx + 1) : 0.0f;
  float s = __fadd_rn(__fmul_rn(v0, w0), __fmul_rn(v1, w1));
  s = __fadd_rn(s, __fmul_rn(v2, w2));
  s = __fadd_rn(s, __fmul_rn(v3, w3));
  return s;
}

__device__ __forceinline__ uint32_t sat_u16_rn(float v) {
  int iv = __float2int_rn(v);
  return (uint32_t)min(max(iv, 0), 65535);
}

// block (64,4): each thread produces 2 horizontally adjacent pixels; grid (ceil(W/128), ceil(H/4), frames)
__global__ void __launch_bounds__(256)
k_warp_affine_u16(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int W, int H,
                  const int* __restrict__ tab, int interp, int skip_frame) {
  const int f = blockIdx.z;
  const int x = (blockIdx.x * 64 + threadIdx.x) * 2;
  const int y = blockIdx.y * 4 + threadIdx.y;
  if (x >= W || y >= H) return;
  const size_t P = (size_t)W * H;
  const uint16_t* s = src + (size_t)f * P;
  uint16_t* d = dst + (size_t)f * P + (size_t)y * W + x;
  if (f == skip_frame) {  // global frame 0 is never registered (psp_process.cpp:1777)
    d[0] = s[(size_t)y * W + x];
    if (x + 1 < W) d[1] = s[(size_t)y * W + x + 1];
    return;
  }
  const int* t = tab + (size_t)f * (2 * W + 2 * H);
  const int X0 = t[2 * W + y], Y0 = t[2 * W + H + y];
  uint32_t o[2] = {0, 0};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    int xx = x + k;
    if (xx >= W) break;
    int X = X0 + t[xx], Y = Y0 + t[W + xx];
    if (interp == 0) {
      int sx = X >> 10, sy = Y >> 10;
      o[k] = (sx >= 0 && sx < W && sy >= 0 && sy < H) ? (uint32_t)__ldg(s + (size_t)sy * W + sx) : 0u;
    } else {
      o[k] = sat_u16_rn(warp_sample_linear<uint16_t>(s, W, H, X, Y));
    }
  }
  if (x + 1 < W && ((W & 1) == 0))
    *reinterpret_cast<uint32_t*>(d) = o[0] | (o[1] << 16);
  else {
    d[0] = (uint16_t)o[0];
    if (x + 1 < W) d[1] = (uint16_t)o[1];
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

__global__ void __launch_bounds__(32)
k_patch(PatchGeom g, const int* __restrict__ cl_list, const uint16_t* __restrict__ frames,
        size_t npix, int nframes, int bstride, float* __restrict__ scratch,
        float* __restrict__ pv) {
  const int cl = cl_list[blockIdx.x];
  const int b = blockIdx.y * 32 + threadIdx.x;
  if (b >= nframes) return;
  const int nz = g.nzp[cl];
  if (nz == 0) return;
  const int off = g.bounds_off[cl], nb = g.bounds_off[cl + 1] - off;
  const uint16_t* img = frames + (size_t)b * npix;
  float* c = scratch + (size_t)off * bstride + b;  // c[i] at c[i*bstride]
  for (int i = 0; i < nb; ++i) {
    int s = __ldg(g.bsrc + off + i);
    c[(size_t)i * bstride] = s >= 0 ? (float)img[s] : pv[(size_t)(-1 - s) * bstride + b];
  }
  const float* E = g.qr_e + (size_t)10 * off;
  const float* TE = g.qr_te + (size_t)10 * off;
  const float* hc = g.hcoef + cl * 10;
  for (int k = 0; k < nz; ++k) {
    const int n = nb - k;
    const float tau = hc[k];
    if (n == 1) {
      c[(size_t)k * bstride] = __fmul_rn(c[(size_t)k * bstride], __fsub_rn(1.0f, tau));
    } else if (tau != 0.0f) {
      const float* e = E + (size_t)k * nb + k + 1;
      const float* te = TE + (size_t)k * nb + k + 1;
      float s = 0.0f;
      for (int i = 0; i < n - 1; ++i)
        s = __fadd_rn(s, __fmul_rn(__ldg(e + i), c[(size_t)(k + 1 + i) * bstride]));
      const float t = __fadd_rn(s, c[(size_t)k * bstride]);
      c[(size_t)k * bstride] = __fsub_rn(c[(size_t)k * bstride], __fmul_rn(tau, t));
      for (int i = 0; i < n - 1; ++i) {
        float* ci = c + (size_t)(k + 1 + i) * bstride;
        *ci = __fsub_rn(*ci, __fmul_rn(t, __ldg(te + i)));
      }
    }
  }
  float x[10], poly[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) x[i] = i < nz ? c[(size_t)i * bstride] : 0.0f;
  for (int i = nz - 1; i >= 0; --i) {
    x[i] = __fdiv_rn(x[i], __ldg(E + (size_t)i * nb + i));
    for (int j = 0; j < i; ++j) x[j] = __fsub_rn(x[j], __fmul_rn(x[i], __ldg(E + (size_t)i * nb + j)));
  }
  const int* pm = g.perm + cl * 10;
#pragma unroll
  for (int i = 0; i < 10; ++i) poly[i] = 0.0f;
  for (int i = 0; i < 10; ++i) {
    float v = i < nz ? x[i] : 0.0f;
    int pi = pm[i];
#pragma unroll
    for (int q = 0; q < 10; ++q)
      if (q == pi) poly[q] = v;
  }
  // polyval2D: z = sum_c poly[c] * y^i * x^j, order [1,x,x2,x3,y,xy,x2y,y2,xy2,y3]
  const int ioff = g.internal_off[cl], ni = g.internal_off[cl + 1] - ioff;
  for (int i = 0; i < ni; ++i) {
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
          z = __fadd_rn(z, __fmul_rn(__fmul_rn(poly[cnt], yp[a]), xp[bb]));
          ++cnt;
        }
    pv[(size_t)(ioff + i) * bstride + b] = z;
  }
}

}  // namespace upsp
