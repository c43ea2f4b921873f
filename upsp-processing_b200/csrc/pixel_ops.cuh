// pixel_ops.cuh -- inline device functions shared by the translation units (upsp_gpu.cu, proj_tma.cu):
// 12-bit unpack, hot-pixel bookkeeping, cv::warpAffine sampling at single pixels.
#pragma once
#include "common.cuh"

namespace upsp {

__device__ __forceinline__ void note_hot(uint32_t v, size_t pix, int thresh, int* cnt, int* pos) {
  if ((int)v >= thresh) {
    int s = atomicAdd(cnt, 1);
    if (s < UPSP_HOT_STORE) pos[s] = (int)pix;
  }
}

__device__ __forceinline__ void unpack12_x8(uint32_t w0, uint32_t w1, uint32_t w2, uint4& o) {
  const uint32_t v0 = __byte_perm(w0, w1, 0x1201);
  const uint32_t v1 = __byte_perm(w0, w1, 0x4534);
  const uint32_t v2 = __byte_perm(w1, w2, 0x3423);
  const uint32_t v3 = __byte_perm(w2, w2, 0x2312);
  o.x = ((v0 >> 4) & 0x00000FFFu) | (v0 & 0x0FFF0000u);
  o.y = ((v1 >> 4) & 0x00000FFFu) | (v1 & 0x0FFF0000u);
  o.z = ((v2 >> 4) & 0x00000FFFu) | (v2 & 0x0FFF0000u);
  o.w = ((v3 >> 4) & 0x00000FFFu) | (v3 & 0x0FFF0000u);
}

template <typename LoadT>
__device__ __forceinline__ float warp_sample_linear(const LoadT* __restrict__ src, int W, int H,
                                                    int X, int Y) {
  X >>= 5;
  Y >>= 5;
  const int sx = X >> 5, sy = Y >> 5;
  const float fx = frac32_exact(X & 31), fy = frac32_exact(Y & 31);
  if (sx >= W || sx + 1 < 0 || sy >= H || sy + 1 < 0) return 0.0f;
  const float w0 = __fmul_rn(1.0f - fy, 1.0f - fx), w1 = __fmul_rn(1.0f - fy, fx);
  const float w2 = __fmul_rn(fy, 1.0f - fx), w3 = __fmul_rn(fy, fx);
  const bool x0 = sx >= 0, x1 = sx + 1 < W, y0 = sy >= 0, y1 = sy + 1 < H;
  const LoadT* r0 = src + (size_t)(y0 ? sy : 0) * W;
  const LoadT* r1 = src + (size_t)(y1 ? sy + 1 : 0) * W;
  float v0 = (x0 && y0) ? u2f_exact(__ldg(r0 + sx)) : 0.0f;
  float v1 = (x1 && y0) ? u2f_exact(__ldg(r0 + sx + 1)) : 0.0f;
  float v2 = (x0 && y1) ? u2f_exact(__ldg(r1 + sx)) : 0.0f;
  float v3 = (x1 && y1) ? u2f_exact(__ldg(r1 + sx + 1)) : 0.0f;
  float s = __fadd_rn(__fmul_rn(v0, w0), __fmul_rn(v1, w1));
  s = __fadd_rn(s, __fmul_rn(v2, w2));
  s = __fadd_rn(s, __fmul_rn(v3, w3));
  return s;
}

// border / nearest-neighbour pixels: rare, kept out of the hot loop's code
static __device__ __noinline__ float warp_px_slow(const uint16_t* __restrict__ s, int W, int H, int X, int Y,
                                           int interp) {
  if (interp == 0) {
    const int sx = X >> 10, sy = Y >> 10;
    return ((unsigned)sx < (unsigned)W && (unsigned)sy < (unsigned)H) ? (float)s[(size_t)sy * W + sx] : 0.0f;
  }
  const float v = warp_sample_linear<uint16_t>(s, W, H, X, Y);
  const float r = __fadd_rn(__fadd_rn(v, 12582912.0f), -12582912.0f);
  return fminf(fmaxf(r, 0.0f), 65535.0f);
}

// hot-pixel fixes of one frame as a list: the pixels fix_hot_pixels (cpp/utils/cv_extras.cpp:230-272) actually
// replaces, with their new values (packed-source mode: the decoded frame is never materialised)
struct HotFix {
  int n;                        // 0..5
  int pos[UPSP_HOT_MAX];
  int val[UPSP_HOT_MAX];
  int pad;
};

// one pixel of a packed 12-bit frame (3 bytes = 2 px, MSB first: cpp/lib/PSPVideo.cpp:134-150)
__device__ __forceinline__ uint32_t px_packed12(const uint8_t* __restrict__ fr, unsigned idx) {
  const uint8_t* p = fr + (size_t)(idx >> 1) * 3;
  return (idx & 1) ? (((uint32_t)__ldg(p + 1) & 0xFu) << 8) | __ldg(p + 2) : ((uint32_t)__ldg(p) << 4) | (__ldg(p + 1) >> 4);
}
__device__ __forceinline__ uint32_t px_packed12_fixed(const uint8_t* __restrict__ fr, unsigned idx, const HotFix* __restrict__ h) {
  if (h != nullptr) {
    const int n = h->n;
    for (int i = 0; i < n; ++i)
      if (h->pos[i] == (int)idx) return (uint32_t)h->val[i];
  }
  return px_packed12(fr, idx);
}

}  // namespace upsp
