// kernels_project.cuh -- K5: batched-frame CSR projection (pixel -> node gather) with the
// camera blend, NaN fill, overlap remap, sum / sum-of-squares and row store fused in.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "pixel_ops.cuh"
#include "project_args.cuh"

namespace upsp {

// Reference: upsp::project_frame cpp/lib/projection.ipp:884-908 (Eigen row-major CSR x
// dense vector) + cpp/exec/psp_process.cpp:1814-1839 (camera sum in camera order, NaN for
// skipped nodes, double-precision sum / sum-sq, adjust_solution, row store).
//
// The overlap remap out[n] = sol[src[n]] is folded into the tables at setup (row n of the
// device table IS row src[n] of the caller's CSR), so the kernel has no remap step; the
// statistics of a remapped node then equal those of its source node, which is what the
// reference's finals produce after adjust_solution(avg/rms) (psp_process.cpp:1936-1939).
//
// Pixel codes: >= 0 pixel index into the camera's u16 frame; <= -2 patched-pixel slot
// (-2 - slot) into the f32 patch-value table (patched pixels are f32 in the reference,
// patches.ipp:104-108,159); -1 (ELL-1 table only) "no entry for this camera".
__device__ __forceinline__ float fetch_px(const ProjCam& c, int code, int b, int bstride) {
  if (code >= 0)
    return c.frames32 ? __ldg(c.frames32 + (size_t)b * c.npix + code)
                      : u2f_exact(__ldg(c.frames + (size_t)b * c.npix + code));
  return __ldg(c.pv + (size_t)(-2 - code) * bstride + b);
}

// ---- fast path: every (remapped) row has <= 1 entry per camera: the reference's own case
// (one nearest pixel per node, psp_process.cpp:318-322).  Thread = node, loop over the
// batch's frames; writes of a warp are 128 contiguous bytes of one intensity row.
template <int NC, int UNROLL>
__global__ void __launch_bounds__(256)
k_project_ell1(const ProjArgs a) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.n_nodes) return;
  int code[NC];
  float val[NC];
  bool skipped = true;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    code[c] = __ldg(a.cam[c].code + n);
    val[c] = __ldg(a.cam[c].val + n);
    skipped = skipped && (code[c] == -1);
  }
  double s = 0.0, q = 0.0;
  float* out = a.out + n;
  for (int b0 = 0; b0 < a.nframes; b0 += UNROLL) {
    float sol[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int b = b0 + u;
      sol[u] = 0.0f;
      if (b < a.nframes) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float cs = 0.0f;
          if (code[c] != -1)
            cs = __fadd_rn(0.0f, __fmul_rn(val[c], fetch_px(a.cam[c], code[c], b, a.bstride)));
          sol[u] = (c == 0) ? cs : __fadd_rn(sol[u], cs);
        }
        if (skipped) sol[u] = __int_as_float(0x7fc00000);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int b = b0 + u;
      if (b < a.nframes) {
        out[(size_t)b * a.n_nodes] = sol[u];
        q += (double)__fmul_rn(sol[u], sol[u]);
        s += (double)sol[u];
      }
    }
  }
  a.sum[n] += s;
  a.sumsq[n] += q;
}

// ---- general CSR path (any number of entries per row; cfg-5's nnz/row 4 and 9 variants).
// Thread = node walking its (short) row; one accumulator per row per camera, entries in CSR
// order, exactly Eigen's row-major sparse * dense loop.
template <int UNROLL>
__global__ void __launch_bounds__(256)
k_project_csr(const ProjArgs a) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.n_nodes) return;
  bool skipped = true;
  for (int c = 0; c < a.n_cams; ++c)
    skipped = skipped && (__ldg(a.cam[c].rowptr + n) == __ldg(a.cam[c].rowptr + n + 1));
  double s = 0.0, q = 0.0;
  float* out = a.out + n;
  for (int b0 = 0; b0 < a.nframes; b0 += UNROLL) {
    float sol[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) sol[u] = 0.0f;
    for (int c = 0; c < a.n_cams; ++c) {
      const ProjCam& cam = a.cam[c];
      const int k0 = __ldg(cam.rowptr + n), k1 = __ldg(cam.rowptr + n + 1);
      float t[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) t[u] = 0.0f;
      for (int k = k0; k < k1; ++k) {
        const int code = __ldg(cam.code + k);
        const float v = __ldg(cam.val + k);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (b0 + u < a.nframes)
            t[u] = __fadd_rn(t[u], __fmul_rn(v, fetch_px(cam, code, b0 + u, a.bstride)));
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        float cs = __fadd_rn(0.0f, t[u]);
        sol[u] = (c == 0) ? cs : __fadd_rn(sol[u], cs);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int b = b0 + u;
      if (b < a.nframes) {
        float v = skipped ? __int_as_float(0x7fc00000) : sol[u];
        out[(size_t)b * a.n_nodes] = v;
        q += (double)__fmul_rn(v, v);
        s += (double)v;
      }
    }
  }
  a.sum[n] += s;
  a.sumsq[n] += q;
}

// ---- fused path (ELL-1 tables): register + project + blend + NaN + sum/sum-sq + remap AND the
// frame-major -> node-major transpose + all-to-all in ONE kernel.
//   * registration: only the pixels the nodes look at are warped (warp_px_u16), straight from
//     the decoded frame; the registered image is never written (saves 2P write + 2P read per
//     frame and half of the per-pixel warp arithmetic when N < P);
//   * output: a block owns 256 nodes x 32 frames, stages the [32][256] result tile in shared
//     memory and writes each node's 32 consecutive frames as one 128-byte segment of its
//     node-major row -- in the destination rank's buffer (peer-mapped over NVLink when the node
//     belongs to another GPU).  The frame-major intensity buffer and the separate transpose
//     pass (8N bytes per frame) disappear; the exchange overlaps phase-1 compute.
// Reference: psp_process.cpp:1790-1842 + local_transpose/global_transpose :647-771.
// ---- building blocks of the fused kernel: U consecutive frames of one camera for one node.
// Every global load of a group is issued before any result is consumed (2U table loads, then
// 4U tap loads in flight per thread); with one camera the table loads of group g+1 are issued
// before the taps of group g are consumed, so the two dependent round trips overlap.  The rare
// border / nearest / unregistered-frame pixels are patched up afterwards through the
// out-of-line slow path (one branch per group).
// INT12: all pixels are < 2^14 (12/10-bit containers), so OpenCV's float bilinear sum
// (weights k/1024, every product and partial sum exact in float) equals S/1024 with the integer
// S = sum t_ij * w_ij; the kernel then rounds S half-to-even in integer arithmetic and never
// touches the conversion (XU) pipe.  u16 containers keep the float sequence.
template <int U>
struct Taps {
  unsigned short t00[U], t01[U], t10[U], t11[U];
  bool all_fast;
};

template <int U>
__device__ __forceinline__ void fused_tabs(const FusedCam& cam, const int2* __restrict__ ptx,
                                           const int2* __restrict__ pty, int (&X)[U], int (&Y)[U]) {
  const size_t tstride = (size_t)(cam.W + cam.H);
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int2 xa = __ldg(ptx + j * tstride), ya = __ldg(pty + j * tstride);
    X[j] = ya.x + xa.x;
    Y[j] = ya.y + xa.y;
  }
}

template <int U>
__device__ __forceinline__ void fused_taps(const FusedCam& cam, const uint16_t* __restrict__ fr,
                                           const int (&X)[U], const int (&Y)[U], Taps<U>& t) {
  t.all_fast = true;
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int sx = X[j] >> 10, sy = Y[j] >> 10;
    const bool fast = (unsigned)sx < (unsigned)(cam.W - 1) && (unsigned)sy < (unsigned)(cam.H - 1);
    t.all_fast = t.all_fast && fast;
    const unsigned idx = fast ? (unsigned)(sy * cam.W + sx) : 0u;
    const uint16_t* p = fr + (j * cam.npix + idx);
    const uint16_t* p2 = p + cam.W;
    t.t00[j] = __ldg(p);
    t.t01[j] = __ldg(p + 1);
    t.t10[j] = __ldg(p2);
    t.t11[j] = __ldg(p2 + 1);
  }
}

template <int U, bool INT12>
__device__ __forceinline__ void fused_finish(const FusedCam& cam, int code, const uint16_t* __restrict__ fr,
                                             const int (&X)[U], const int (&Y)[U], const Taps<U>& t, int b,
                                             int interp, int skip_frame, float (&v)[U]) {
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int fxi = (X[j] >> 5) & 31, fyi = (Y[j] >> 5) & 31;
    if (INT12) {
      const int gx = 32 - fxi, gy = 32 - fyi;
      int S = (int)t.t00[j] * (gy * gx);
      S += (int)t.t01[j] * (gy * fxi);
      S += (int)t.t10[j] * (fyi * gx);
      S += (int)t.t11[j] * (fyi * fxi);
      const int qv = S >> 10, rem = S & 1023;
      const int r = qv + ((rem + (qv & 1)) > 512);      // round half to even
      v[j] = u2f_exact((uint32_t)r);
    } else {
      const float fx = frac32_exact(fxi), fy = frac32_exact(fyi);
      const float gx = 1.0f - fx, gy = 1.0f - fy;
      float r = __fadd_rn(__fmul_rn((float)t.t00[j], __fmul_rn(gy, gx)), __fmul_rn((float)t.t01[j], __fmul_rn(gy, fx)));
      r = __fadd_rn(r, __fmul_rn((float)t.t10[j], __fmul_rn(fy, gx)));
      r = __fadd_rn(r, __fmul_rn((float)t.t11[j], __fmul_rn(fy, fx)));
      r = __fadd_rn(__fadd_rn(r, 12582912.0f), -12582912.0f);   // rint (half to even)
      v[j] = fminf(r, 65535.0f);                                 // r >= 0 by construction
    }
  }
  const bool has_skip = (unsigned)(skip_frame - b) < (unsigned)U;
  if (!t.all_fast || interp != 1 || has_skip) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int sx = X[j] >> 10, sy = Y[j] >> 10;
      const bool fast = interp == 1 && (unsigned)sx < (unsigned)(cam.W - 1) && (unsigned)sy < (unsigned)(cam.H - 1);
      const uint16_t* f2 = fr + j * cam.npix;
      if (b + j == skip_frame) v[j] = (float)__ldg(f2 + code);
      else if (!fast) v[j] = warp_px_slow(f2, cam.W, cam.H, X[j], Y[j], interp);
    }
  }
}

template <int U, bool REG, bool INT12>
__device__ __forceinline__ void fused_cam_group(const FusedCam& cam, int code, const int2* __restrict__ ptx,
                                                const int2* __restrict__ pty, const uint16_t* __restrict__ fr,
                                                int b, int interp, int skip_frame, int bstride,
                                                float (&v)[U]) {
  if (code >= 0) {
    if (REG) {
      int X[U], Y[U];
      Taps<U> t;
      fused_tabs<U>(cam, ptx, pty, X, Y);
      fused_taps<U>(cam, fr, X, Y, t);
      fused_finish<U, INT12>(cam, code, fr, X, Y, t, b, interp, skip_frame, v);
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j) v[j] = (float)__ldg(fr + (j * cam.npix + (size_t)code));
    }
  } else if (code <= -2) {
    const float* p = cam.pv + (size_t)(-2 - code) * bstride + b;
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = __ldg(p + j);
  }
}

template <int NC, bool REG, bool INT12, int U, int BS>
__global__ void __launch_bounds__(BS)
k_project_fused(const FusedArgs a) {
  __shared__ float tile[32][BS + 1];
  __shared__ float* rowp[BS];        // node-major row of each of the block's nodes (any rank)
  const int gid = blockIdx.x * BS + threadIdx.x;
  const bool live = gid < a.n_nodes;
  const int n = live ? __ldg(a.perm + gid) : -1;
  {
    rowp[threadIdx.x] = live ? fused_row_ptr(a, n) : nullptr;
  }
  int code[NC];
  const int2* ptx[NC];
  const int2* pty[NC];
  float val[NC];
  bool skipped = true;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    code[c] = live ? __ldg(a.cam[c].code + n) : -1;
    val[c] = live ? __ldg(a.cam[c].val + n) : 0.0f;
    skipped = skipped && (code[c] == -1);
    const int W = a.cam[c].W;
    const int px = code[c] >= 0 ? code[c] % W : 0, py = code[c] >= 0 ? code[c] / W : 0;
    ptx[c] = reinterpret_cast<const int2*>(a.cam[c].tab) + px;        // (adelta,bdelta)[px] of frame 0
    pty[c] = reinterpret_cast<const int2*>(a.cam[c].tab) + W + py;    // (X0,Y0)[py] of frame 0
  }
  double s = 0.0, q = 0.0;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // one camera: software-pipelining the table loads across groups was measured SLOWER (0.47 vs
  // 0.37 ms per 128-frame batch: register pressure), so it stays off
  constexpr bool PIPE = false;
  for (int b0 = 0; b0 < a.nframes; b0 += 32) {
    const int nb = min(32, a.nframes - b0);
    if (live) {
      int u = 0;
      if (PIPE && code[0] >= 0) {
        const FusedCam& cam = a.cam[0];
        const size_t ts = (size_t)(cam.W + cam.H);
        int X[U], Y[U], Xn[U], Yn[U];
        if (U <= nb) fused_tabs<U>(cam, ptx[0] + (size_t)b0 * ts, pty[0] + (size_t)b0 * ts, X, Y);
        for (; u + U <= nb; u += U) {
          const int b = b0 + u;
          const uint16_t* fr = cam.frames + (size_t)b * cam.npix;
          Taps<U> t;
          fused_taps<U>(cam, fr, X, Y, t);
          // tables of the next group (clamped to this one at the end of the chunk: harmless reload)
          const int bn = (u + 2 * U <= nb) ? b + U : b;
          fused_tabs<U>(cam, ptx[0] + (size_t)bn * ts, pty[0] + (size_t)bn * ts, Xn, Yn);
          float v[U];
          fused_finish<U, INT12>(cam, code[0], fr, X, Y, t, b, a.interp, a.skip_frame, v);
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const float sol = __fadd_rn(0.0f, __fmul_rn(val[0], v[j]));
            tile[u + j][threadIdx.x] = sol;
            q += (double)__fmul_rn(sol, sol);
            s += (double)sol;
            X[j] = Xn[j];
            Y[j] = Yn[j];
          }
        }
      }
      for (; u + U <= nb; u += U) {
        float sol[U];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float v[U];
#pragma unroll
          for (int j = 0; j < U; ++j) v[j] = 0.0f;
          const FusedCam& cam = a.cam[c];
          const int b = b0 + u;
          fused_cam_group<U, REG, INT12>(cam, code[c], ptx[c] + (size_t)b * (cam.W + cam.H),
                                         pty[c] + (size_t)b * (cam.W + cam.H),
                                         cam.frames + (size_t)b * cam.npix, b, a.interp, a.skip_frame,
                                         a.bstride, v);
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const float cs = (code[c] == -1) ? 0.0f : __fadd_rn(0.0f, __fmul_rn(val[c], v[j]));
            sol[j] = (c == 0) ? cs : __fadd_rn(sol[j], cs);
          }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
          if (skipped) sol[j] = __int_as_float(0x7fc00000);
          tile[u + j][threadIdx.x] = sol[j];
          q += (double)__fmul_rn(sol[j], sol[j]);
          s += (double)sol[j];
        }
      }
      for (; u < nb; ++u) {   // tail frames of the batch, one at a time
        float sol = 0.0f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float v[1] = {0.0f};
          const FusedCam& cam = a.cam[c];
          const int b = b0 + u;
          fused_cam_group<1, REG, INT12>(cam, code[c], ptx[c] + (size_t)b * (cam.W + cam.H),
                                         pty[c] + (size_t)b * (cam.W + cam.H),
                                         cam.frames + (size_t)b * cam.npix, b, a.interp, a.skip_frame,
                                         a.bstride, v);
          const float cs = (code[c] == -1) ? 0.0f : __fadd_rn(0.0f, __fmul_rn(val[c], v[0]));
          sol = (c == 0) ? cs : __fadd_rn(sol, cs);
        }
        if (skipped) sol = __int_as_float(0x7fc00000);
        tile[u][threadIdx.x] = sol;
        q += (double)__fmul_rn(sol, sol);
        s += (double)sol;
      }
    }
    __syncthreads();
    // warp w writes the block's nodes [w*32, w*32+32): lane = frame -> 128-byte row segments
    if (lane < nb) {
#pragma unroll 4
      for (int j = 0; j < 32; ++j) {
        float* rp = rowp[w * 32 + j];
        if (rp == nullptr) break;
        rp[b0 + lane] = tile[lane][w * 32 + j];
      }
    }
    __syncthreads();
  }
  if (live) {
    a.sum[n] += s;
    a.sumsq[n] += q;
  }
}

// ---- Lean fused kernels for the hot configuration (registration on, bilinear, pixels < 2^13, i.e.
// 12-bit containers).  Same arithmetic as k_project_fused, same results bit for bit; what changed is
// the cost per node-frame:
//   * every address is base + 32-bit element offset (one IMAD.WIDE per load instead of 64-bit
//     add chains); frame strides of the batch are warp-uniform;
//   * the bilinear sum S = sum t_ij w_ij (integer, < 2^22) is accumulated on top of the float
//     bit pattern of 2^23, so S never needs an int->float conversion: the float 2^23 + S times
//     2^-10 plus (1.5 * 2^23 - 2^13) lands in [2^23, 2^24) where one FFMA rounds S/1024 half to
//     even (OpenCV's saturate_cast<ushort>(float) = cvRound), and one FADD removes the offset;
//   * val * v + 0 is one FFMA (x*y + 0.0 rounds once, exactly like the FMUL + FADD pair, and
//     turns -0 into +0 the same way);
//   * a thread stores its 4 frames of a group with one STS.128 into a [node][frame] tile, and the
//     write-out moves 4 frames per lane (LDS.128 + STG.128).
// Preconditions (host-checked): batch * npix < 2^31, registration tables present, interp linear.
template <int U>
struct Taps2 {
  unsigned t00[U], t01[U], t10[U], t11[U];
};

// The warp coordinates of a node-frame are NOT gathered from the per-frame tables (measured: 1.3 GB
// of L2->L1 traffic per 128-frame batch, ~1 GB of it the (adelta,bdelta)[x] table that every image
// row's nodes re-read, and two dependent global round trips per group):
//   * (adelta, bdelta)[x] = cvRound(M0*x*1024), cvRound(M3*x*1024) are evaluated in the thread
//     (DMUL + F2I.S32.F64, exactly k_warp_tables' expression: M*1024 is an exact scaling) from the
//     per-frame coefficient pairs staged in shared memory;
//   * (X0, Y0)[y]: a block's nodes are consecutive in raster order, so they span a few image rows;
//     the block copies those rows' entries into shared memory and the per-node-frame lookup is one
//     LDS.64.  Blocks whose nodes span more than RY rows of a camera fall back to the global y-table.
// The tap loads are then the only global loads of a group: one round trip instead of two.
constexpr int FUSED3_RY = 4;

template <int U, bool CHK = true>
__device__ __forceinline__ void fused3_cam_group(const FusedCam& cam, int code, double dpx, const int2* __restrict__ ysrc,
                                                 unsigned ystep, const double2* __restrict__ coef, unsigned b,
                                                 int skip_frame, float (&v)[U]) {
  const uint16_t* __restrict__ fr = cam.frames;
  const unsigned W = (unsigned)cam.W, npix = (unsigned)cam.npix;
  int X[U], Y[U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const double2 cf = coef[j];
    const int2 ya = ysrc[j * ystep];
    X[j] = ya.x + __double2int_rn(__dmul_rn(cf.x, dpx));
    Y[j] = ya.y + __double2int_rn(__dmul_rn(cf.y, dpx));
  }
  Taps2<U> t;
  bool all_fast = true;
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int sx = X[j] >> 10, sy = Y[j] >> 10;
    // CHK = false: the block has proven (fused4, per frame, from the corners of its pixel box) that
    // every tap of every one of its nodes lies inside the image
    const bool fast = !CHK || ((unsigned)sx < (unsigned)(cam.W - 1) && (unsigned)sy < (unsigned)(cam.H - 1));
    all_fast = all_fast && fast;
    const unsigned idx = (fast ? (unsigned)(sy * cam.W + sx) : 0u) + (b + j) * npix;
    const uint16_t* p0 = fr + idx;
    const uint16_t* p1 = fr + (idx + W);
    t.t00[j] = __ldg(p0);
    t.t01[j] = __ldg(p0 + 1);
    t.t10[j] = __ldg(p1);
    t.t11[j] = __ldg(p1 + 1);
  }
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const unsigned fxi = ((unsigned)X[j] >> 5) & 31u, fyi = ((unsigned)Y[j] >> 5) & 31u;
    const unsigned gx = 32u - fxi, gy = 32u - fyi;
    const unsigned top = t.t00[j] * gx + t.t01[j] * fxi;
    const unsigned bot = t.t10[j] * gx + t.t11[j] * fxi;
    const unsigned sm = top * gy + 0x4B000000u + bot * fyi;          // bits of the float 2^23 + S
    v[j] = __fadd_rn(__fmaf_rn(__uint_as_float(sm), 0.0009765625f, 12574720.0f), -12582912.0f);
  }
  const bool has_skip = CHK && (unsigned)(skip_frame - (int)b) < (unsigned)U;
  if (CHK && (!all_fast || has_skip)) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int sx = X[j] >> 10, sy = Y[j] >> 10;
      const bool fast = (unsigned)sx < (unsigned)(cam.W - 1) && (unsigned)sy < (unsigned)(cam.H - 1);
      const uint16_t* f2 = fr + (size_t)(b + j) * cam.npix;
      if ((int)b + j == skip_frame) v[j] = (float)__ldg(f2 + code);
      else if (!fast) v[j] = warp_px_slow(f2, cam.W, cam.H, X[j], Y[j], 1);
    }
  }
}

// ---- k_project_fused4: the fused kernel of the hot configuration, with warp-private output tiles.
// A warp computes 32 nodes and writes out the same 32 nodes, so the [node][frame] staging tile is
// private to the warp and the compute -> write-out hand-over needs __syncwarp only: the main loop
// has no block barrier (with a block-wide tile 8 % of the warp time sat in BAR, and a block ran at
// the pace of its slowest warp).  The per-frame tables (coefficient pairs, y-table rows) are staged once per STAGE
// frames (the whole 128-frame batch for one camera) instead of once per 32-frame chunk.  Chunks are
// 16 frames (64-byte row segments, two chunks fill a 128-byte line back to back), which halves the
// tile: shared memory per block drops from 21 KB to 17 KB and the unified L1 keeps ~30 KB more for
// the tap lines.
// CH = 32 (128-byte row segments) is used when rows go straight into peer memory: NVLink write
// efficiency halves with 64-byte segments (measured at 8 GPUs: 0.73 vs 0.40 ms per 128 frames).
template <int NC, int BS, int CH = 16, int GU = 4>
__global__ void __launch_bounds__(BS, 1024 / BS)
k_project_fused4(const FusedArgs a) {
  constexpr int RY = FUSED3_RY;
  constexpr int S = NC == 1 ? 128 : NC == 2 ? 64 : 32;      // frames per table stage
  constexpr int TS = CH + 4;      // tile row stride in floats (frames + pad; STS.128 conflict-free for 20 and 36)
  static_assert(CH == 16 || CH == 32, "chunk of 16 or 32 frames");
  __shared__ __align__(16) float tile[BS * TS];        // [node][frame], rows [w*32, w*32+32) private to warp w
  __shared__ float* rowp[BS];
  __shared__ __align__(16) double2 s_coef[NC][S];      // (M0, M3) * 1024 of the stage's frames
  __shared__ __align__(8) int2 s_y[NC][S * RY];        // (X0, Y0)[y0 .. y0+RY) of the stage's frames
  __shared__ int s_ymin[NC], s_ymax[NC];
  __shared__ int s_xmin[NC], s_xmax[NC];               // column range of the block's node pixels
  __shared__ __align__(4) unsigned char s_fast[NC][S]; // per frame: all taps of all the block's nodes are interior
  const int gid = blockIdx.x * BS + threadIdx.x;
  const bool live = gid < a.n_nodes;
  const int n = live ? __ldg(a.perm + gid) : -1;
  if (threadIdx.x < NC) {
    s_ymin[threadIdx.x] = 0x7fffffff;
    s_ymax[threadIdx.x] = -1;
    s_xmin[threadIdx.x] = 0x7fffffff;
    s_xmax[threadIdx.x] = -1;
  }
  {
    rowp[threadIdx.x] = live ? fused_row_ptr(a, n) : nullptr;
  }
  __syncthreads();
  int code[NC];
  float val[NC];
  double dpx[NC];
  int py[NC];
  bool skipped = true;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    code[c] = live ? __ldg(a.cam[c].code + n) : -1;
    val[c] = live ? __ldg(a.cam[c].val + n) : 0.0f;
    skipped = skipped && (code[c] == -1);
    const int W = a.cam[c].W;
    dpx[c] = code[c] >= 0 ? (double)(code[c] % W) : 0.0;
    py[c] = code[c] >= 0 ? code[c] / W : 0;
    if (code[c] >= 0) {
      atomicMin(&s_ymin[c], py[c]);
      atomicMax(&s_ymax[c], py[c]);
      atomicMin(&s_xmin[c], code[c] % W);
      atomicMax(&s_xmax[c], code[c] % W);
    }
  }
  __syncthreads();
  int y0[NC];
  bool ysm[NC];
  unsigned ystep[NC];          // distance (in int2) between consecutive frames' y entries
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    y0[c] = s_ymin[c];
    ysm[c] = s_ymax[c] - y0[c] < RY;     // also true when the block has no plain-pixel node of this camera
    ystep[c] = ysm[c] ? (unsigned)RY : (unsigned)(a.cam[c].W + a.cam[c].H);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0, q = 0.0;
  const bool vec_ok = ((a.f_total | a.col0) & 3) == 0;
  float* trow = tile + threadIdx.x * TS;

  for (int s0 = 0; s0 < a.nframes; s0 += S) {
    const int ns = min(S, a.nframes - s0);
    if (s0 > 0) __syncthreads();       // every warp is done with the previous stage's tables
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const FusedCam& cam = a.cam[c];
      for (int i = threadIdx.x; i < S; i += BS) {
        bool fast = false;
        if (i < ns) {
          const float* M = cam.m6 + (size_t)(s0 + i) * 6;
          const double2 cf = make_double2((double)__ldg(M) * 1024.0, (double)__ldg(M + 3) * 1024.0);
          s_coef[c][i] = cf;
          // X(x,y) = cvRound(M0 x 1024) + X0[y] is monotone in x and in y (same for Y), so over the
          // block's pixel box it is extremal at the four corners
          if (s_ymax[c] >= 0 && s0 + i != a.skip_frame) {
            const int2* tab2 = reinterpret_cast<const int2*>(cam.tab);
            const unsigned tb = (unsigned)(s0 + i) * (unsigned)(cam.W + cam.H) + (unsigned)cam.W;
            const int2 ylo = __ldg(tab2 + (tb + (unsigned)s_ymin[c])), yhi = __ldg(tab2 + (tb + (unsigned)s_ymax[c]));
            const double xl = (double)s_xmin[c], xh = (double)s_xmax[c];
            const int axl = __double2int_rn(cf.x * xl), axh = __double2int_rn(cf.x * xh);
            const int bxl = __double2int_rn(cf.y * xl), bxh = __double2int_rn(cf.y * xh);
            fast = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int sx = (((k & 1) ? axh : axl) + ((k & 2) ? yhi.x : ylo.x)) >> 10;
              const int sy = (((k & 1) ? bxh : bxl) + ((k & 2) ? yhi.y : ylo.y)) >> 10;
              fast = fast && (unsigned)sx < (unsigned)(cam.W - 1) && (unsigned)sy < (unsigned)(cam.H - 1);
            }
          }
        }
        s_fast[c][i] = fast ? 1 : 0;
      }
      if (ysm[c] && s_ymax[c] >= 0) {
        const int2* tab2 = reinterpret_cast<const int2*>(cam.tab);
        const unsigned ts = (unsigned)(cam.W + cam.H);
        for (int i = threadIdx.x; i < ns * RY; i += BS) {
          const int f = s0 + i / RY, r = i % RY;
          const int yy = min(y0[c] + r, cam.H - 1);
          s_y[c][i] = __ldg(tab2 + ((unsigned)f * ts + (unsigned)(cam.W + yy)));
        }
      }
    }
    __syncthreads();
    for (int c0 = 0; c0 < ns; c0 += CH) {
      const int nb = min(CH, ns - c0);
      const int b0 = s0 + c0;
      if (live) {
        const int2* ysrc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c)
          ysrc[c] = ysm[c] ? &s_y[c][c0 * RY + (py[c] - y0[c])]
                           : reinterpret_cast<const int2*>(a.cam[c].tab) +
                                 ((unsigned)b0 * ystep[c] + (unsigned)(a.cam[c].W + py[c]));
        auto group = [&](auto UU, int u, float* sol) {
          constexpr int U = decltype(UU)::value;
          if constexpr (NC == 1) {
            if (code[0] >= 0) {
              float v[U];
              bool nochk = false;
              if constexpr (U == 4) nochk = *reinterpret_cast<const unsigned*>(&s_fast[0][c0 + u]) == 0x01010101u;
              if (nochk)
                fused3_cam_group<U, false>(a.cam[0], code[0], dpx[0], ysrc[0] + (unsigned)u * ystep[0], ystep[0],
                                           &s_coef[0][c0 + u], (unsigned)(b0 + u), a.skip_frame, v);
              else
                fused3_cam_group<U>(a.cam[0], code[0], dpx[0], ysrc[0] + (unsigned)u * ystep[0], ystep[0],
                                    &s_coef[0][c0 + u], (unsigned)(b0 + u), a.skip_frame, v);
#pragma unroll
              for (int j = 0; j < U; ++j) sol[j] = __fmaf_rn(val[0], v[j], 0.0f);
            } else if (code[0] <= -2) {
              const float* p = a.cam[0].pv + (size_t)(-2 - code[0]) * a.bstride + (b0 + u);
#pragma unroll
              for (int j = 0; j < U; ++j) sol[j] = __fmaf_rn(val[0], __ldg(p + j), 0.0f);
            } else {
#pragma unroll
              for (int j = 0; j < U; ++j) sol[j] = __int_as_float(0x7fc00000);
            }
          } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              float v[U];
#pragma unroll
              for (int j = 0; j < U; ++j) v[j] = 0.0f;
              const FusedCam& cam = a.cam[c];
              if (code[c] >= 0) {
                bool nochk = false;
                if constexpr (U == 4) nochk = *reinterpret_cast<const unsigned*>(&s_fast[c][c0 + u]) == 0x01010101u;
                if (nochk)
                  fused3_cam_group<U, false>(cam, code[c], dpx[c], ysrc[c] + (unsigned)u * ystep[c], ystep[c],
                                             &s_coef[c][c0 + u], (unsigned)(b0 + u), a.skip_frame, v);
                else
                  fused3_cam_group<U>(cam, code[c], dpx[c], ysrc[c] + (unsigned)u * ystep[c], ystep[c],
                                      &s_coef[c][c0 + u], (unsigned)(b0 + u), a.skip_frame, v);
              } else if (code[c] <= -2) {
                const float* p = cam.pv + (size_t)(-2 - code[c]) * a.bstride + (b0 + u);
#pragma unroll
                for (int j = 0; j < U; ++j) v[j] = __ldg(p + j);
              }
#pragma unroll
              for (int j = 0; j < U; ++j) {
                const float cs = (code[c] == -1) ? 0.0f : __fmaf_rn(val[c], v[j], 0.0f);
                sol[j] = (c == 0) ? cs : __fadd_rn(sol[j], cs);
              }
            }
            if (skipped) {
#pragma unroll
              for (int j = 0; j < U; ++j) sol[j] = __int_as_float(0x7fc00000);
            }
          }
        };
        int u = 0;
        for (; u + GU <= nb; u += GU) {
          float sol[GU];
          group(std::integral_constant<int, GU>{}, u, sol);
          if constexpr (GU == 4) *reinterpret_cast<float4*>(trow + u) = make_float4(sol[0], sol[1], sol[2], sol[3]);
          else if constexpr (GU == 2) *reinterpret_cast<float2*>(trow + u) = make_float2(sol[0], sol[1]);
          else {
#pragma unroll
            for (int j = 0; j < GU; ++j) trow[u + j] = sol[j];
          }
#pragma unroll
          for (int j = 0; j < GU; ++j) {
            q += (double)__fmul_rn(sol[j], sol[j]);
            s += (double)sol[j];
          }
        }
        for (; u < nb; ++u) {
          float sol[1];
          group(std::integral_constant<int, 1>{}, u, sol);
          trow[u] = sol[0];
          q += (double)__fmul_rn(sol[0], sol[0]);
          s += (double)sol[0];
        }
      }
      __syncwarp();
      if (vec_ok && nb == CH) {
        // LPN lanes = one node's CH frames (64 or 128 bytes); 32 / LPN nodes per pass
        constexpr int LPN = CH / 4;
        const int fq = (lane % LPN) * 4;
#pragma unroll
        for (int it = 0; it < LPN; ++it) {
          const int nl = w * 32 + it * (32 / LPN) + lane / LPN;
          float* rp = rowp[nl];
          if (rp != nullptr) {
            const float4 o = *reinterpret_cast<const float4*>(tile + nl * TS + fq);
            *reinterpret_cast<float4*>(rp + b0 + fq) = o;
          }
        }
      } else if (lane < nb) {
        for (int j = 0; j < 32; ++j) {
          float* rp = rowp[w * 32 + j];
          if (rp != nullptr) rp[b0 + lane] = tile[(w * 32 + j) * TS + lane];
        }
      }
      __syncwarp();
    }
  }
  if (live) {
    a.sum[n] += s;
    a.sumsq[n] += q;
  }
}

// stand-alone project_frame on f32 frames (upsp_op_project_frames): out[f][r]
__global__ void __launch_bounds__(256)
k_project_f32(const int* __restrict__ rowptr, const int* __restrict__ col,
              const float* __restrict__ val, int n_rows, const float* __restrict__ frames,
              size_t npix, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (r >= n_rows) return;
  const float* fr = frames + (size_t)f * npix;
  float t = 0.0f;
  for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) t = __fadd_rn(t, __fmul_rn(val[k], __ldg(fr + col[k])));
  out[(size_t)f * n_rows + r] = __fadd_rn(0.0f, t);
}

// a10: finals (cpp/exec/psp_process.cpp:1930-1935): avg = (float)(sum/F), rms = (float)sqrt(sumsq/F)
__global__ void k_phase1_finals(const double* __restrict__ sum, const double* __restrict__ sumsq,
                                int n, unsigned n_frames_total, float* __restrict__ avg,
                                float* __restrict__ rms) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  avg[i] = (float)(sum[i] / (double)n_frames_total);
  rms[i] = (float)sqrt(sumsq[i] / (double)n_frames_total);
}

}  // namespace upsp
