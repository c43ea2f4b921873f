// kernels_ecc.cuh -- K2: batched affine ECC registration (cv::findTransformECC as called by
// upsp::register_pixel, cpp/lib/registration.cpp:43-64: MOTION_AFFINE, no mask,
// gaussFiltSize 5, criteria COUNT+EPS 50 / 1e-3).
#pragma once
#include "common.cuh"

namespace upsp {

// OpenCV's algorithm (video/src/ecc.cpp), restated:
//   T = GaussianBlur(template, 5x5, sigma 0)  -> separable [1 4 6 4 1]/16, BORDER_REFLECT_101
//   I = GaussianBlur(frame as f32, 5x5);  gx = (I[x+1]-I[x-1])/2, gy likewise (REFLECT_101)
//   per iteration with the current 2x3 map M (identity at start):
//     Iw, gxw, gyw = warpAffine(I | gx | gy, M, INTER_LINEAR | WARP_INVERSE_MAP)  (fixed-point model)
//     mask         = warpAffine(ones, M, INTER_NEAREST | WARP_INVERSE_MAP)
//     masked means / norms of Iw and T; J = [gxw X, gyw X, gxw Y, gyw Y, gxw, gyw];
//     H = J^T J; rho = <T~, I~> / (|T~||I~|); ip = J^T I~; tp = J^T T~;
//     lambda = (|I~|^2 - ip.H^-1 ip) / (<T~,I~> - tp.H^-1 ip); dp = H^-1 (lambda tp - ip); M += dp
//   until |rho - rho_prev| < eps or 50 iterations.
// Here every image-wide sum of one iteration is produced by ONE pass over the pixels
// (k_ecc_reduce: 42 raw moments per frame; the zero-mean forms follow algebraically), and a
// one-thread-per-frame kernel (k_ecc_solve) does the 6x6 algebra in double and updates M.

constexpr int ECC_NSUM = 42;   // 18 Hessian + 6 sum J + 6 sum J*Iw + 6 sum J*T + 6 stats
constexpr int ECC_NT = 256;
constexpr int ECC_BATCH = 32;  // frames solved together (working set: 16 B/px/frame)
constexpr int ECC_NBLK = 64;   // row strips per frame in k_ecc_reduce

__device__ __forceinline__ int reflect101(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}

// row pass of the 5-tap blur; SRC = uint16_t (frame) or float
template <typename SRC>
__global__ void __launch_bounds__(256)
k_ecc_blur_rows(const SRC* __restrict__ src, float* __restrict__ dst, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const size_t f = blockIdx.z;
  if (x >= W) return;
  const SRC* r = src + (f * H + y) * (size_t)W;
  const float k0 = 0.0625f, k1 = 0.25f, k2 = 0.375f;
  float s = __fmul_rn(k0, (float)r[reflect101(x - 2, W)]);
  s = __fadd_rn(s, __fmul_rn(k1, (float)r[reflect101(x - 1, W)]));
  s = __fadd_rn(s, __fmul_rn(k2, (float)r[x]));
  s = __fadd_rn(s, __fmul_rn(k1, (float)r[reflect101(x + 1, W)]));
  s = __fadd_rn(s, __fmul_rn(k0, (float)r[reflect101(x + 2, W)]));
  dst[(f * H + y) * (size_t)W + x] = s;
}

__global__ void __launch_bounds__(256)
k_ecc_blur_cols(const float* __restrict__ src, float* __restrict__ dst, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const size_t f = blockIdx.z;
  if (x >= W) return;
  const float* b = src + f * H * (size_t)W + x;
  const float k0 = 0.0625f, k1 = 0.25f, k2 = 0.375f;
  float s = __fmul_rn(k0, b[(size_t)reflect101(y - 2, H) * W]);
  s = __fadd_rn(s, __fmul_rn(k1, b[(size_t)reflect101(y - 1, H) * W]));
  s = __fadd_rn(s, __fmul_rn(k2, b[(size_t)y * W]));
  s = __fadd_rn(s, __fmul_rn(k1, b[(size_t)reflect101(y + 1, H) * W]));
  s = __fadd_rn(s, __fmul_rn(k0, b[(size_t)reflect101(y + 2, H) * W]));
  dst[(f * H + y) * (size_t)W + x] = s;
}

// gx, gy interleaved as float2 so the warp of both gradients is one 8-byte tap
__global__ void __launch_bounds__(256)
k_ecc_grad(const float* __restrict__ img, float2* __restrict__ g, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const size_t f = blockIdx.z;
  if (x >= W) return;
  const float* b = img + f * H * (size_t)W;
  const float gx = __fmul_rn(0.5f, __fsub_rn(b[(size_t)y * W + reflect101(x + 1, W)], b[(size_t)y * W + reflect101(x - 1, W)]));
  const float gy = __fmul_rn(0.5f, __fsub_rn(b[(size_t)reflect101(y + 1, H) * W + x], b[(size_t)reflect101(y - 1, H) * W + x]));
  g[(f * H + y) * (size_t)W + x] = make_float2(gx, gy);
}

struct EccState {          // per local frame
  float rho, last_rho;
  int iters;
  int status;              // 0 running, 1 converged / iteration cap, 2 failed (NaN rho or lambda_d <= 0), 3 skipped
};

// One pass over one frame's pixels for one iteration: 42 raw moments.
//   grid (NBLK, B); block b handles rows [b*rows_per_block, ...); thread = column strip
//   tab: this iteration's fixed-point warp tables of the batch ([(ad,bd)[W] | (X0,Y0)[H]], linear)
__global__ void __launch_bounds__(ECC_NT)
k_ecc_reduce(const float* __restrict__ I, const float2* __restrict__ G, const float* __restrict__ T,
             const int* __restrict__ tab, const EccState* __restrict__ st, int W, int H,
             int rows_per_block, double* __restrict__ partial /* [B][NBLK][ECC_NSUM] */) {
  __shared__ float park[ECC_NSUM * ECC_NT];
  const int f = blockIdx.y;
  if (st[f].status != 0) return;
  const float* If = I + (size_t)f * W * H;
  const float2* Gf = G + (size_t)f * W * H;
  const int2* t2 = reinterpret_cast<const int2*>(tab + (size_t)f * (2 * W + 2 * H));
  float acc[ECC_NSUM];
#pragma unroll
  for (int k = 0; k < ECC_NSUM; ++k) acc[k] = 0.0f;
  const int y0 = blockIdx.x * rows_per_block, y1 = min(H, y0 + rows_per_block);
  for (int x = threadIdx.x; x < W; x += ECC_NT) {
    const int2 xa = __ldg(t2 + x);
    const float Xf = (float)x;
    for (int y = y0; y < y1; ++y) {
      const int2 ya = __ldg(t2 + W + y);
      const int X = ya.x + xa.x, Y = ya.y + xa.y;          // round_delta 16 included
      // mask: INTER_NEAREST uses round_delta 512 and >> 10
      const int nx = (X + 496) >> 10, ny = (Y + 496) >> 10;
      const bool m = (unsigned)nx < (unsigned)W && (unsigned)ny < (unsigned)H;
      const int Xs = X >> 5, Ys = Y >> 5;
      const int sx = Xs >> 5, sy = Ys >> 5;
      float iw = 0.0f, gx = 0.0f, gy = 0.0f;
      if (!(sx >= W || sx + 1 < 0 || sy >= H || sy + 1 < 0)) {
        const float fx = frac32_exact(Xs & 31), fy = frac32_exact(Ys & 31);
        const float w00 = __fmul_rn(1.0f - fy, 1.0f - fx), w01 = __fmul_rn(1.0f - fy, fx);
        const float w10 = __fmul_rn(fy, 1.0f - fx), w11 = __fmul_rn(fy, fx);
        const bool x0 = sx >= 0, x1 = sx + 1 < W, yy0 = sy >= 0, yy1 = sy + 1 < H;
        const size_t r0 = (size_t)(yy0 ? sy : 0) * W, r1 = (size_t)(yy1 ? sy + 1 : 0) * W;
        const int c0 = x0 ? sx : 0, c1 = x1 ? sx + 1 : 0;
        const bool v00 = x0 && yy0, v01 = x1 && yy0, v10 = x0 && yy1, v11 = x1 && yy1;
        const float i00 = v00 ? __ldg(If + r0 + c0) : 0.0f, i01 = v01 ? __ldg(If + r0 + c1) : 0.0f;
        const float i10 = v10 ? __ldg(If + r1 + c0) : 0.0f, i11 = v11 ? __ldg(If + r1 + c1) : 0.0f;
        const float2 z = make_float2(0.0f, 0.0f);
        const float2 g00 = v00 ? __ldg(Gf + r0 + c0) : z, g01 = v01 ? __ldg(Gf + r0 + c1) : z;
        const float2 g10 = v10 ? __ldg(Gf + r1 + c0) : z, g11 = v11 ? __ldg(Gf + r1 + c1) : z;
        iw = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(i00, w00), __fmul_rn(i01, w01)), __fmul_rn(i10, w10)), __fmul_rn(i11, w11));
        gx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(g00.x, w00), __fmul_rn(g01.x, w01)), __fmul_rn(g10.x, w10)), __fmul_rn(g11.x, w11));
        gy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(g00.y, w00), __fmul_rn(g01.y, w01)), __fmul_rn(g10.y, w10)), __fmul_rn(g11.y, w11));
      }
      const float Yf = (float)y;
      const float tm = m ? __ldg(T + (size_t)y * W + x) : 0.0f;
      const float im = m ? iw : 0.0f;
      // Hessian moments {a,b,c} x {X2, XY, Y2, X, Y, 1}
      const float a = gx * gx, b = gx * gy, c = gy * gy;
      const float mono[6] = {Xf * Xf, Xf * Yf, Yf * Yf, Xf, Yf, 1.0f};
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        acc[k] = fmaf(a, mono[k], acc[k]);
        acc[6 + k] = fmaf(b, mono[k], acc[6 + k]);
        acc[12 + k] = fmaf(c, mono[k], acc[12 + k]);
      }
      // J = [gx X, gy X, gx Y, gy Y, gx, gy]
      const float J[6] = {gx * Xf, gy * Xf, gx * Yf, gy * Yf, gx, gy};
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        acc[18 + k] += m ? J[k] : 0.0f;              // mean terms act under the mask only
        acc[24 + k] = fmaf(J[k], iw, acc[24 + k]);   // I~ = Iw outside the mask (subtract is masked)
        acc[30 + k] = fmaf(J[k], tm, acc[30 + k]);
      }
      acc[36] += m ? 1.0f : 0.0f;
      acc[37] += im;
      acc[38] = fmaf(im, im, acc[38]);
      acc[39] += tm;
      acc[40] = fmaf(tm, tm, acc[40]);
      acc[41] = fmaf(tm, im, acc[41]);
    }
  }
  // block sum -> double partials (fixed order)
#pragma unroll
  for (int k = 0; k < ECC_NSUM; ++k) park[k * ECC_NT + threadIdx.x] = acc[k];
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = w; k < ECC_NSUM; k += ECC_NT / 32) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < ECC_NT / 32; ++i) t += (double)park[k * ECC_NT + i * 32 + lane];
    t = warp_sum(t);
    if (lane == 0) partial[((size_t)f * gridDim.x + blockIdx.x) * ECC_NSUM + k] = t;
  }
}

// 6x6 inverse in double, Gauss-Jordan with partial pivoting; returns false if singular
__device__ inline bool inv6(const double (&A)[6][6], double (&R)[6][6]) {
  double a[6][12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      a[i][j] = A[i][j];
      a[i][6 + j] = i == j ? 1.0 : 0.0;
    }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r)
      if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
    if (a[p][c] == 0.0) return false;
    if (p != c)
      for (int j = 0; j < 12; ++j) {
        const double t = a[p][j];
        a[p][j] = a[c][j];
        a[c][j] = t;
      }
    const double d = 1.0 / a[c][c];
    for (int j = 0; j < 12; ++j) a[c][j] *= d;
    for (int r = 0; r < 6; ++r) {
      if (r == c) continue;
      const double m = a[r][c];
      if (m == 0.0) continue;
      for (int j = 0; j < 12; ++j) a[r][j] -= m * a[c][j];
    }
  }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) R[i][j] = a[i][6 + j];
  return true;
}

// One thread per frame: finish the iteration (ecc.cpp main loop body) and decide about the next.
__global__ void k_ecc_solve(const double* __restrict__ partial, int nblk, int nframes, float* __restrict__ m6,
                            EccState* __restrict__ st, int max_iters, float eps, int* __restrict__ n_active) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nframes) return;
  EccState s = st[f];
  if (s.status != 0) return;
  double S[ECC_NSUM];
  for (int k = 0; k < ECC_NSUM; ++k) {
    double t = 0.0;
    for (int b = 0; b < nblk; ++b) t += partial[((size_t)f * nblk + b) * ECC_NSUM + k];
    S[k] = t;
  }
  const double *a = S, *b = S + 6, *c = S + 12;   // x {X2, XY, Y2, X, Y, 1}
  double Hd[6][6] = {
      {a[0], b[0], a[1], b[1], a[3], b[3]},
      {b[0], c[0], b[1], c[1], b[3], c[3]},
      {a[1], b[1], a[2], b[2], a[4], b[4]},
      {b[1], c[1], b[2], c[2], b[4], c[4]},
      {a[3], b[3], a[4], b[4], a[5], b[5]},
      {b[3], c[3], b[4], c[4], b[5], c[5]}};
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) Hd[i][j] = (double)(float)Hd[i][j];   // OpenCV keeps H as CV_32F
  const double cnt = S[36];
  const double mI = S[37] / cnt, mT = S[39] / cnt;
  const double in2 = fmax(S[38] / cnt - mI * mI, 0.0) * cnt;   // |I~|^2 = cnt * std^2
  const double tn2 = fmax(S[40] / cnt - mT * mT, 0.0) * cnt;
  const double corr = S[41] - cnt * mT * mI;
  double ip[6], tp[6];
  for (int k = 0; k < 6; ++k) {
    ip[k] = (double)(float)(S[24 + k] - mI * S[18 + k]);
    tp[k] = (double)(float)(S[30 + k] - mT * S[18 + k]);
  }
  s.iters += 1;
  s.last_rho = s.rho;
  const double rho = corr / (sqrt(in2) * sqrt(tn2));
  s.rho = (float)rho;
  double Hi[6][6];
  bool ok = !(rho != rho) && inv6(Hd, Hi);
  double lam_d = 0.0, lam_n = 0.0, iph[6];
  if (ok) {
    for (int i = 0; i < 6; ++i) {
      double t = 0.0;
      for (int j = 0; j < 6; ++j) t += (double)(float)Hi[i][j] * ip[j];
      iph[i] = (double)(float)t;
    }
    double d1 = 0.0, d2 = 0.0;
    for (int i = 0; i < 6; ++i) {
      d1 += ip[i] * iph[i];
      d2 += tp[i] * iph[i];
    }
    lam_n = in2 - d1;
    lam_d = corr - d2;
    ok = lam_d > 0.0;
  }
  if (!ok) {
    s.status = 2;   // the reference throws cv::Exception here
    st[f] = s;
    return;
  }
  const double lam = lam_n / lam_d;
  double ep[6], dp[6];
  for (int k = 0; k < 6; ++k) ep[k] = (double)(float)(lam * tp[k] - ip[k]);
  for (int i = 0; i < 6; ++i) {
    double t = 0.0;
    for (int j = 0; j < 6; ++j) t += (double)(float)Hi[i][j] * ep[j];
    dp[i] = t;
  }
  float* M = m6 + (size_t)f * 6;   // update_warping_matrix_ECC, MOTION_AFFINE
  M[0] += (float)dp[0];
  M[3] += (float)dp[1];
  M[1] += (float)dp[2];
  M[4] += (float)dp[3];
  M[2] += (float)dp[4];
  M[5] += (float)dp[5];
  // loop condition of the NEXT iteration: (i <= N) && (fabs(rho - last_rho) >= eps)
  if (s.iters >= max_iters || fabs((double)s.rho - (double)s.last_rho) < (double)eps) s.status = 1;
  else atomicAdd(n_active, 1);
  st[f] = s;
}

__global__ void k_ecc_init(EccState* __restrict__ st, float* __restrict__ m6, int nframes, int skip_frame,
                           float eps) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nframes) return;
  EccState s;
  s.rho = -1.0f;
  s.last_rho = -eps;
  s.iters = 0;
  s.status = f == skip_frame ? 3 : 0;
  st[f] = s;
  float* M = m6 + (size_t)f * 6;
  M[0] = 1.0f; M[1] = 0.0f; M[2] = 0.0f;
  M[3] = 0.0f; M[4] = 1.0f; M[5] = 0.0f;
}

}  // namespace upsp
