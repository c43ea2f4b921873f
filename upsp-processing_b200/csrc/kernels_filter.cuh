// kernels_filter.cuh -- K4: the optional spatial filter between patching and projection
// (cpp/exec/psp_process.cpp:1802-1807): cv::GaussianBlur(img, img, Size(k,k), 0) or
// cv::blur(img, img, Size(k,k)), default border (BORDER_REFLECT_101).  The image is CV_32F when
// the polynomial patcher ran (patches.ipp:104-108) and CV_16U otherwise, and OpenCV treats the
// two depths differently:
//   * CV_16U Gaussian: fixed-point kernel (16 fractional bits per pass), exact integer
//     accumulation, one rounding at the end -- reproduced exactly (pinned against cv2 goldens);
//   * CV_16U box: integer window sum, times 1/area in double, cvRound;
//   * CV_32F Gaussian: separable float filter (exact in any summation order on 12-bit data);
//   * CV_32F box: window sum in double, times 1/area, narrowed to float.
// CV_16U Gaussian: any odd size 3..31 (taps from gauss_fixed.inc = OpenCV's error-diffusion rounding of its bit-exact
// sigma = 0 kernel).  CV_32F Gaussian (patcher on): k = 3, 5, 7 only, whose taps are dyadic so that every product is
// exact; for larger sizes the float result depends on OpenCV's summation order inside its SIMD row / column filters
// (measured: no ordering of the separable sum reproduces cv2 4.13 within 1 ulp), so those are rejected.
#pragma once
#include "common.cuh"
#include "kernels_ecc.cuh"   // reflect101

namespace upsp {

struct FilterSpec {
  int kind;        // 1 gaussian, 2 box
  int ksize;       // odd
  int kq[31];      // gaussian taps in 16.16 fixed point (gauss_fixed.inc)
  float kf[7];     // gaussian taps as float, k <= 7
};

__global__ void __launch_bounds__(256)
k_filter_u16(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int W, int H, FilterSpec fs) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const size_t f = blockIdx.z;
  if (x >= W) return;
  const uint16_t* img = src + f * (size_t)W * H;
  const int r = fs.ksize / 2;
  if (fs.kind == 1) {
    long long S = 0;
    for (int j = -r; j <= r; ++j) {
      const uint16_t* row = img + (size_t)reflect101(y + j, H) * W;
      long long t = 0;
      for (int i = -r; i <= r; ++i) t += (long long)fs.kq[i + r] * (long long)row[reflect101(x + i, W)];
      S += (long long)fs.kq[j + r] * t;
    }
    const long long v = (S + (1LL << 31)) >> 32;
    dst[f * (size_t)W * H + (size_t)y * W + x] = (uint16_t)min(max(v, 0LL), 65535LL);
  } else {
    int S = 0;
    for (int j = -r; j <= r; ++j) {
      const uint16_t* row = img + (size_t)reflect101(y + j, H) * W;
      for (int i = -r; i <= r; ++i) S += row[reflect101(x + i, W)];
    }
    const int v = __double2int_rn(__dmul_rn((double)S, 1.0 / (double)(fs.ksize * fs.ksize)));
    dst[f * (size_t)W * H + (size_t)y * W + x] = (uint16_t)min(max(v, 0), 65535);
  }
}

// f32: pass = 0 row pass (src -> dst), pass = 1 column pass; box does everything in pass 0
__global__ void __launch_bounds__(256)
k_filter_f32(const float* __restrict__ src, float* __restrict__ dst, int W, int H, FilterSpec fs, int pass) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const size_t f = blockIdx.z;
  if (x >= W) return;
  const float* img = src + f * (size_t)W * H;
  const int r = fs.ksize / 2;
  float out;
  if (fs.kind == 1) {
    // symmetric form s = c*k0 + sum_i (a[-i] + a[+i]) * k_i  (OpenCV's SymmRowSmallVec / SymmColumnVec)
    if (pass == 0) {
      const float* row = img + (size_t)y * W;
      out = __fmul_rn(row[x], fs.kf[r]);
      for (int i = 1; i <= r; ++i)
        out = __fadd_rn(out, __fmul_rn(__fadd_rn(row[reflect101(x - i, W)], row[reflect101(x + i, W)]), fs.kf[r + i]));
    } else {
      out = __fmul_rn(img[(size_t)y * W + x], fs.kf[r]);
      for (int i = 1; i <= r; ++i)
        out = __fadd_rn(out, __fmul_rn(__fadd_rn(img[(size_t)reflect101(y - i, H) * W + x],
                                                 img[(size_t)reflect101(y + i, H) * W + x]), fs.kf[r + i]));
    }
  } else {
    double S = 0.0;
    for (int j = -r; j <= r; ++j) {
      const float* row = img + (size_t)reflect101(y + j, H) * W;
      double t = 0.0;
      for (int i = -r; i <= r; ++i) t += (double)row[reflect101(x + i, W)];
      S += t;
    }
    out = (float)__dmul_rn(S, 1.0 / (double)(fs.ksize * fs.ksize));
  }
  dst[f * (size_t)W * H + (size_t)y * W + x] = out;
}

// u16 frame -> f32 image with the patched interior pixels (f32) substituted: the CV_32F image
// PatchClusters::operator() returns.  Two launches: convert, then scatter the final slots.
__global__ void __launch_bounds__(256)
k_u16_to_f32(const uint16_t* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}
__global__ void __launch_bounds__(256)
k_scatter_patched(const float* __restrict__ pv, const int* __restrict__ slot_pix /* -1: superseded */,
                  int n_slots, int bstride, int nframes, size_t npix, float* __restrict__ img) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (s >= n_slots || b >= nframes) return;
  const int pix = slot_pix[s];
  if (pix >= 0) img[(size_t)b * npix + pix] = pv[(size_t)s * bstride + b];
}

}  // namespace upsp
