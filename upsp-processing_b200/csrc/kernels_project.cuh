// kernels_project.cuh -- K5: batched-frame CSR projection (pixel -> node gather) with the
// camera blend, NaN fill, overlap remap, sum / sum-of-squares and row store fused in.
#pragma once
#include <type_traits>

#include "common.cuh"

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
struct ProjCam {
  const uint16_t* frames;  // [batch][npix] registered u16 frames of this batch
  const float* frames32;   // or (filter after patching) the f32 image of the batch; overrides `frames`
  size_t npix;
  const float* pv;         // [slots][bstride] patch values of this batch (or nullptr)
  const int* code;         // ELL-1: [N]; CSR: [nnz]
  const float* val;
  const int* rowptr;       // CSR only: [N+1]
};
struct ProjArgs {
  int n_cams, n_nodes, nframes, bstride;
  ProjCam cam[UPSP_MAX_CAMS];
  float* out;      // first row of this batch in the frame-major intensity buffer [F][N]
  double* sum;     // [N]  += over the batch
  double* sumsq;   // [N]
};

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
struct FusedCam {
  const uint16_t* frames;  // [batch][npix] decoded, hot-pixel-fixed frames (NOT registered)
  size_t npix;
  int W, H;
  const int* tab;          // [batch][2W+2H] warp tables, or nullptr (registration = none)
  const float* m6;         // [batch][6] the 2x3 maps the tables were built from (k_project_fused3)
  const float* pv;         // [slots][bstride] patch values
  const int* code;         // [N]
  const float* val;        // [N]
};
struct FusedArgs {
  int n_cams, n_nodes, nframes, bstride, interp, skip_frame;
  int dbg;                             // experiment switches (UPSP_FUSED_DBG): 1 no row stores, 2 taps from one hot line, 4 no statistics
  FusedCam cam[UPSP_MAX_CAMS];
  double* sum;
  double* sumsq;
  const int* perm;                     // [N] processing order: nodes sorted by pixel index (raster),
                                       // so a warp gathers from one or two image rows
  int perm_len;                        // staged kernel: padded length of its (tile-ordered) perm
  int n_ranks, f_total, col0;          // col0 = global frame index of the batch's first frame
  float* dst[UPSP_MAX_RANKS];          // node-major [N_s][F] buffer of every rank
  int node_start[UPSP_MAX_RANKS + 1];
};

// border / nearest-neighbour pixels: rare, kept out of the hot loop's code
__device__ __noinline__ float warp_px_slow(const uint16_t* __restrict__ s, int W, int H, int X, int Y,
                                           int interp) {
  if (interp == 0) {
    const int sx = X >> 10, sy = Y >> 10;
    return ((unsigned)sx < (unsigned)W && (unsigned)sy < (unsigned)H) ? (float)s[(size_t)sy * W + sx] : 0.0f;
  }
  const float v = warp_sample_linear<uint16_t>(s, W, H, X, Y);
  const float r = __fadd_rn(__fadd_rn(v, 12582912.0f), -12582912.0f);
  return fminf(fmaxf(r, 0.0f), 65535.0f);
}

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
    float* rp = nullptr;
    if (live) {
      int r = 0;
      while (r + 1 < a.n_ranks && n >= a.node_start[r + 1]) ++r;
      rp = a.dst[r] + (size_t)(n - a.node_start[r]) * a.f_total + a.col0;
    }
    rowp[threadIdx.x] = rp;
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

// ---- k_project_fused2: the lean version of the fused kernel for the hot configuration
// (registration on, bilinear, pixels < 2^13, i.e. 12-bit containers).  Same arithmetic, same
// results bit for bit; what changed is the instruction budget per node-frame (87 -> ~45):
//   * every address is base + 32-bit element offset (one IMAD.WIDE per load instead of 64-bit
//     add chains); frame / table strides of the batch are warp-uniform;
//   * the bilinear sum S = sum t_ij w_ij (integer, < 2^22) is accumulated on top of the float
//     bit pattern of 2^23, so S never needs an int->float conversion: the float 2^23 + S times
//     2^-10 plus (1.5 * 2^23 - 2^13) lands in [2^23, 2^24) where one FFMA rounds S/1024 half to
//     even (OpenCV's saturate_cast<ushort>(float) = cvRound), and one FADD removes the offset;
//   * val * v + 0 is one FFMA (x*y + 0.0 rounds once, exactly like the FMUL + FADD pair, and
//     turns -0 into +0 the same way);
//   * a thread stores its 4 frames of a group with one STS.128 into a [node][36] tile, and the
//     write-out moves 4 frames per lane (LDS.128 + STG.128, 8 lanes = one 128-byte row segment).
// Preconditions (host-checked): batch * npix < 2^31, registration tables present, interp linear.
template <int U>
struct Taps2 {
  unsigned t00[U], t01[U], t10[U], t11[U];
};

template <int U>
__device__ __forceinline__ void fused2_cam_group(const FusedCam& cam, int code, unsigned px, unsigned pyw,
                                                 unsigned b, int skip_frame, float (&v)[U]) {
  const int2* __restrict__ tab2 = reinterpret_cast<const int2*>(cam.tab);
  const uint16_t* __restrict__ fr = cam.frames;
  const unsigned ts = (unsigned)(cam.W + cam.H), W = (unsigned)cam.W, npix = (unsigned)cam.npix;
  int X[U], Y[U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const unsigned tb = (b + j) * ts;
    const int2 xa = __ldg(tab2 + (tb + px)), ya = __ldg(tab2 + (tb + pyw));
    X[j] = ya.x + xa.x;
    Y[j] = ya.y + xa.y;
  }
  Taps2<U> t;
  bool all_fast = true;
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int sx = X[j] >> 10, sy = Y[j] >> 10;
    const bool fast = (unsigned)sx < (unsigned)(cam.W - 1) && (unsigned)sy < (unsigned)(cam.H - 1);
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
  const bool has_skip = (unsigned)(skip_frame - (int)b) < (unsigned)U;
  if (!all_fast || has_skip) {
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

template <int NC, int U>
__device__ __forceinline__ void fused2_group(const FusedArgs& a, const int (&code)[NC], const float (&val)[NC],
                                             const unsigned (&px)[NC], const unsigned (&pyw)[NC],
                                             bool skipped, unsigned b, float (&sol)[U]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = 0.0f;
    const FusedCam& cam = a.cam[c];
    if (code[c] >= 0) {
      fused2_cam_group<U>(cam, code[c], px[c], pyw[c], b, a.skip_frame, v);
    } else if (code[c] <= -2) {
      const float* p = cam.pv + (size_t)(-2 - code[c]) * a.bstride + b;
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

template <int NC, int BS>
__global__ void __launch_bounds__(BS)
k_project_fused2(const FusedArgs a) {
  constexpr int TS = 36;                              // tile row stride in floats (32 frames + pad, 16-byte aligned)
  __shared__ __align__(16) float tile[BS * TS];       // [node][frame]
  __shared__ float* rowp[BS];
  const int gid = blockIdx.x * BS + threadIdx.x;
  const bool live = gid < a.n_nodes;
  const int n = live ? __ldg(a.perm + gid) : -1;
  {
    float* rp = nullptr;
    if (live) {
      int r = 0;
      while (r + 1 < a.n_ranks && n >= a.node_start[r + 1]) ++r;
      rp = a.dst[r] + (size_t)(n - a.node_start[r]) * a.f_total + a.col0;
    }
    rowp[threadIdx.x] = rp;
  }
  int code[NC];
  float val[NC];
  unsigned px[NC], pyw[NC];
  bool skipped = true;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    code[c] = live ? __ldg(a.cam[c].code + n) : -1;
    val[c] = live ? __ldg(a.cam[c].val + n) : 0.0f;
    skipped = skipped && (code[c] == -1);
    const int W = a.cam[c].W;
    px[c] = code[c] >= 0 ? (unsigned)(code[c] % W) : 0u;
    pyw[c] = (unsigned)W + (code[c] >= 0 ? (unsigned)(code[c] / W) : 0u);
  }
  double s = 0.0, q = 0.0;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec_ok = ((a.f_total | a.col0) & 3) == 0;
  float* trow = tile + threadIdx.x * TS;
  for (int b0 = 0; b0 < a.nframes; b0 += 32) {
    const int nb = min(32, a.nframes - b0);
    if (live) {
      int u = 0;
      for (; u + 4 <= nb; u += 4) {
        float sol[4];
        fused2_group<NC, 4>(a, code, val, px, pyw, skipped, (unsigned)(b0 + u), sol);
        *reinterpret_cast<float4*>(trow + u) = make_float4(sol[0], sol[1], sol[2], sol[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          q += (double)__fmul_rn(sol[j], sol[j]);
          s += (double)sol[j];
        }
      }
      for (; u < nb; ++u) {
        float sol[1];
        fused2_group<NC, 1>(a, code, val, px, pyw, skipped, (unsigned)(b0 + u), sol);
        trow[u] = sol[0];
        q += (double)__fmul_rn(sol[0], sol[0]);
        s += (double)sol[0];
      }
    }
    __syncthreads();
    if (vec_ok && nb == 32) {
      // warp w writes the block's nodes [w*32, w*32+32): 8 lanes = one node's 32 frames = 128 bytes
      const int fq = (lane & 7) * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int nl = w * 32 + it * 4 + (lane >> 3);
        float* rp = rowp[nl];
        if (rp != nullptr) {
          const float4 o = *reinterpret_cast<const float4*>(tile + nl * TS + fq);
          *reinterpret_cast<float4*>(rp + b0 + fq) = o;
        }
      }
    } else if (lane < nb) {
      for (int j = 0; j < 32; ++j) {
        float* rp = rowp[w * 32 + j];
        if (rp == nullptr) break;
        rp[b0 + lane] = tile[(w * 32 + j) * TS + lane];
      }
    }
    __syncthreads();
  }
  if (live) {
    a.sum[n] += s;
    a.sumsq[n] += q;
  }
}

// ---- k_project_fused3: k_project_fused2 without the per-node-frame table gathers.
// ncu on fused2: 1.3 GB of L2->L1 traffic per 128-frame batch, ~1 GB of it the (adelta,bdelta)[x]
// table (every image row's nodes re-read the 8 KB x-table of every frame), two dependent global
// round trips per group (table -> taps), issue 38 %, L1 wavefronts 43 %: latency bound.  Here
//   * (adelta, bdelta)[x] = cvRound(M0*x*1024), cvRound(M3*x*1024) are evaluated in the thread
//     (DMUL + F2I.S32.F64, exactly k_warp_tables' expression: M*1024 is an exact scaling) from the
//     32 per-frame coefficient pairs of the chunk, staged in shared memory;
//   * (X0, Y0)[y]: a block's nodes are consecutive in raster order, so they span a few image rows;
//     the block copies those rows' entries for the chunk's 32 frames into shared memory (while the
//     previous chunk is being written out) and the per-node-frame lookup is one LDS.64.  Blocks
//     whose nodes span more than RY rows of a camera fall back to the global y-table.
// The tap loads are then the only global loads of a group: one round trip instead of two.
constexpr int FUSED3_RY = 4;

template <int U>
__device__ __forceinline__ void fused3_cam_group(const FusedCam& cam, int code, double dpx, const int2* __restrict__ ysrc,
                                                 unsigned ystep, const double2* __restrict__ coef, unsigned b,
                                                 int skip_frame, int dbg, float (&v)[U]) {
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
    const bool fast = (unsigned)sx < (unsigned)(cam.W - 1) && (unsigned)sy < (unsigned)(cam.H - 1);
    all_fast = all_fast && fast;
    const unsigned idx = (dbg & 2) ? (unsigned)(sx & 63) : (fast ? (unsigned)(sy * cam.W + sx) : 0u) + (b + j) * npix;
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
  const bool has_skip = (unsigned)(skip_frame - (int)b) < (unsigned)U;
  if (!all_fast || has_skip) {
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

template <int NC, int BS, int MINB = 1024 / BS>
__global__ void __launch_bounds__(BS, MINB)
k_project_fused3(const FusedArgs a) {
  constexpr int TS = 36;                              // tile row stride in floats (32 frames + pad)
  constexpr int RY = FUSED3_RY;
  __shared__ __align__(16) float tile[BS * TS];       // [node][frame]
  __shared__ float* rowp[BS];
  __shared__ __align__(16) double2 s_coef[NC][32];    // (M0, M3) * 1024 of the chunk's frames
  __shared__ __align__(8) int2 s_y[NC][32 * RY];      // (X0, Y0)[y0 .. y0+RY) of the chunk's frames
  __shared__ int s_ymin[NC], s_ymax[NC];
  const int gid = blockIdx.x * BS + threadIdx.x;
  const bool live = gid < a.n_nodes;
  const int n = live ? __ldg(a.perm + gid) : -1;
  if (threadIdx.x < NC) {
    s_ymin[threadIdx.x] = 0x7fffffff;
    s_ymax[threadIdx.x] = -1;
  }
  {
    float* rp = nullptr;
    if (live) {
      int r = 0;
      while (r + 1 < a.n_ranks && n >= a.node_start[r + 1]) ++r;
      rp = a.dst[r] + (size_t)(n - a.node_start[r]) * a.f_total + a.col0;
    }
    rowp[threadIdx.x] = rp;
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
    }
  }
  __syncthreads();
  int y0[NC];
  bool ysm[NC];
  const int2* ysrc[NC];        // this thread's (X0,Y0) entry of the chunk's first frame
  unsigned ystep[NC];          // distance (in int2) between consecutive frames' entries
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    y0[c] = s_ymin[c];
    ysm[c] = s_ymax[c] - y0[c] < RY;     // also true when the block has no plain-pixel node of this camera
    ystep[c] = ysm[c] ? (unsigned)RY : (unsigned)(a.cam[c].W + a.cam[c].H);
  }
  // stage the coefficient pairs and the y-table rows of frames [b0, b0 + 32)
  auto stage = [&](int b0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const FusedCam& cam = a.cam[c];
      if (threadIdx.x < 32) {
        const int f = min(b0 + (int)threadIdx.x, a.nframes - 1);
        const float* M = cam.m6 + (size_t)f * 6;
        s_coef[c][threadIdx.x] = make_double2((double)__ldg(M) * 1024.0, (double)__ldg(M + 3) * 1024.0);
      }
      if (ysm[c] && s_ymax[c] >= 0) {
        const int2* tab2 = reinterpret_cast<const int2*>(cam.tab);
        const unsigned ts = (unsigned)(cam.W + cam.H);
        for (int i = threadIdx.x; i < 32 * RY; i += BS) {
          const int f = min(b0 + i / RY, a.nframes - 1), r = i % RY;
          const int yy = min(y0[c] + r, cam.H - 1);
          s_y[c][i] = __ldg(tab2 + ((unsigned)f * ts + (unsigned)(cam.W + yy)));
        }
      }
    }
  };
  stage(0);
  __syncthreads();
  double s = 0.0, q = 0.0;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec_ok = ((a.f_total | a.col0) & 3) == 0;
  float* trow = tile + threadIdx.x * TS;
  for (int b0 = 0; b0 < a.nframes; b0 += 32) {
    const int nb = min(32, a.nframes - b0);
    if (live) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        ysrc[c] = ysm[c] ? &s_y[c][py[c] - y0[c]]
                         : reinterpret_cast<const int2*>(a.cam[c].tab) +
                               ((unsigned)b0 * ystep[c] + (unsigned)(a.cam[c].W + py[c]));
      auto group = [&](auto UU, int u, float* sol) {
        constexpr int U = decltype(UU)::value;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float v[U];
#pragma unroll
          for (int j = 0; j < U; ++j) v[j] = 0.0f;
          const FusedCam& cam = a.cam[c];
          if (code[c] >= 0) {
            fused3_cam_group<U>(cam, code[c], dpx[c], ysrc[c] + (unsigned)u * ystep[c], ystep[c], &s_coef[c][u],
                                (unsigned)(b0 + u), a.skip_frame, a.dbg, v);
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
      };
      int u = 0;
      for (; u + 4 <= nb; u += 4) {
        float sol[4];
        group(std::integral_constant<int, 4>{}, u, sol);
        *reinterpret_cast<float4*>(trow + u) = make_float4(sol[0], sol[1], sol[2], sol[3]);
        if (!(a.dbg & 4)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            q += (double)__fmul_rn(sol[j], sol[j]);
            s += (double)sol[j];
          }
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
    __syncthreads();
    if (b0 + 32 < a.nframes) stage(b0 + 32);      // the compute loop is done with this chunk's tables
    if (vec_ok && nb == 32) {
      // warp w writes the block's nodes [w*32, w*32+32): 8 lanes = one node's 32 frames = 128 bytes
      const int fq = (lane & 7) * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int nl = w * 32 + it * 4 + (lane >> 3);
        float* rp = rowp[nl];
        if (rp != nullptr && !(a.dbg & 1)) {
          const float4 o = *reinterpret_cast<const float4*>(tile + nl * TS + fq);
          *reinterpret_cast<float4*>(rp + b0 + fq) = o;
        }
      }
    } else if (lane < nb) {
      for (int j = 0; j < 32; ++j) {
        float* rp = rowp[w * 32 + j];
        if (rp == nullptr) break;
        rp[b0 + lane] = tile[(w * 32 + j) * TS + lane];
      }
    }
    __syncthreads();
  }
  if (live) {
    a.sum[n] += s;
    a.sumsq[n] += q;
  }
}

// ---- staged variant of the fused kernel (one camera, registration on, <= 14-bit pixels).
// The nodes are processed in TILE order (64 x 16 pixel tiles, raster inside a tile, every tile
// padded to whole 128-node blocks), so the pixels a block needs lie in a small rectangle that is
// known at setup time (BlockInfo).  Per stage of `fps` frames the block copies that rectangle
// (+ a registration margin) of the decoded frames, and the matching slices of the warp tables,
// into shared memory with 16-byte cp.async transfers; the per-node work then runs entirely out
// of shared memory (6 LDS instead of 6 scattered LDG per node-frame, no long-scoreboard
// stalls).  Taps that fall outside the staged rectangle (shift larger than the margin, image
// border) take the global slow path -- correctness never depends on the margin.
struct BlockInfo {
  short x0, y0;     // bounding box of the block's node pixels
  short w, h;       // extents (w == 0: no plain-pixel node in this block)
  short tx0, ty0;   // staged rectangle origin (tx0 multiple of 8)
  short tw, th;     // staged extents (tw multiple of 8)
  int fps;          // frames per stage (0: rectangle too large, use the global path)
};
constexpr int STAGE_MARGIN = 3;
constexpr int STAGE_BYTES = 40 * 1024;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <bool INT12>
__global__ void __launch_bounds__(128)
k_project_staged(const FusedArgs a, const BlockInfo* __restrict__ binfo) {
  constexpr int BS = 128;
  extern __shared__ __align__(16) unsigned char stage[];   // [fps][th][tw] u16 | [fps][w] int2 | [fps][h] int2
  __shared__ float tile[32][BS + 1];
  __shared__ float* rowp[BS];
  const FusedCam& cam = a.cam[0];
  const int gid = blockIdx.x * BS + threadIdx.x;
  const int n = gid < a.perm_len ? __ldg(a.perm + gid) : -1;    // perm is padded with -1
  const bool live = n >= 0;
  {
    float* rp = nullptr;
    if (live) {
      int r = 0;
      while (r + 1 < a.n_ranks && n >= a.node_start[r + 1]) ++r;
      rp = a.dst[r] + (size_t)(n - a.node_start[r]) * a.f_total + a.col0;
    }
    rowp[threadIdx.x] = rp;
  }
  const BlockInfo bi = binfo[blockIdx.x];
  const int code = live ? __ldg(cam.code + n) : -1;
  const float val = live ? __ldg(cam.val + n) : 0.0f;
  const int W = cam.W, H = cam.H;
  const int px = code >= 0 ? code % W : 0, py = code >= 0 ? code / W : 0;
  const int tw = bi.tw, th = bi.th, fps = bi.fps;
  const size_t img_bytes = (size_t)tw * th * 2;
  uint16_t* s_img = reinterpret_cast<uint16_t*>(stage);
  int2* s_xa = reinterpret_cast<int2*>(stage + (size_t)fps * img_bytes);
  int2* s_ya = s_xa + (size_t)fps * bi.w;
  const int tstride = W + H;
  double s = 0.0, q = 0.0;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int b0 = 0; b0 < a.nframes; b0 += 32) {
    const int nb = min(32, a.nframes - b0);
    for (int u0 = 0; u0 < nb; u0 += (fps > 0 ? fps : nb)) {
      const int ns = fps > 0 ? min(fps, nb - u0) : nb - u0;
      const int b = b0 + u0;
      if (fps > 0) {
        __syncthreads();   // previous stage fully consumed
        // ---- stage: warp `w` copies frames w, w+4, ... of the stage; lanes walk the rectangle in
        // 16-byte chunks (no integer divisions: the chunk cursor is advanced incrementally)
        const int cpr = tw / 8;                                   // chunks per row
        const int2* tb0 = reinterpret_cast<const int2*>(cam.tab) + (size_t)b * tstride;
        for (int j = w; j < ns; j += BS / 32) {
          const uint16_t* gsrc = cam.frames + (size_t)(b + j) * cam.npix + (size_t)bi.ty0 * W + bi.tx0;
          uint16_t* sdst = s_img + (size_t)j * th * tw;
          int ry = 0, cx = lane;
          while (cx >= cpr) { cx -= cpr; ++ry; }
          while (ry < th) {
            cp_async16(sdst + ry * tw + cx * 8, gsrc + (size_t)ry * W + cx * 8);
            cx += 32;
            while (cx >= cpr) { cx -= cpr; ++ry; }
          }
          const int2* tb = tb0 + (size_t)j * tstride;
          for (int k = lane; k < bi.w; k += 32) cp_async8(s_xa + j * bi.w + k, tb + bi.x0 + k);
          for (int k = lane; k < bi.h; k += 32) cp_async8(s_ya + j * bi.h + k, tb + W + bi.y0 + k);
        }
        cp_async_wait_all();
        __syncthreads();
      }
      if (live) {
        if (code >= 0) {
#pragma unroll 4
          for (int j = 0; j < ns; ++j) {
            const int bj = b + j;
            const uint16_t* fr = cam.frames + (size_t)bj * cam.npix;
            float v;
            if (bj == a.skip_frame) {
              v = (float)__ldg(fr + code);
            } else {
              int2 xa, ya;
              if (fps > 0) {
                xa = s_xa[j * bi.w + (px - bi.x0)];
                ya = s_ya[j * bi.h + (py - bi.y0)];
              } else {
                const int2* tb = reinterpret_cast<const int2*>(cam.tab) + (size_t)bj * tstride;
                xa = __ldg(tb + px);
                ya = __ldg(tb + W + py);
              }
              const int X = ya.x + xa.x, Y = ya.y + xa.y;
              const int sx = X >> 10, sy = Y >> 10;
              const int lx = sx - bi.tx0, ly = sy - bi.ty0;
              const bool in_frame = (unsigned)sx < (unsigned)(W - 1) && (unsigned)sy < (unsigned)(H - 1);
              const bool in_tile = fps > 0 && (unsigned)lx < (unsigned)(tw - 1) && (unsigned)ly < (unsigned)(th - 1);
              if (a.interp == 1 && in_frame && (in_tile || fps == 0)) {
                unsigned t00, t01, t10, t11;
                if (in_tile) {
                  const uint16_t* p = s_img + ((size_t)j * th + ly) * tw + lx;
                  t00 = p[0]; t01 = p[1]; t10 = p[tw]; t11 = p[tw + 1];
                } else {
                  const uint16_t* p = fr + (size_t)sy * W + sx;
                  t00 = __ldg(p); t01 = __ldg(p + 1); t10 = __ldg(p + W); t11 = __ldg(p + W + 1);
                }
                const int fxi = (X >> 5) & 31, fyi = (Y >> 5) & 31;
                if (INT12) {
                  const int gx = 32 - fxi, gy = 32 - fyi;
                  int S = (int)t00 * (gy * gx);
                  S += (int)t01 * (gy * fxi);
                  S += (int)t10 * (fyi * gx);
                  S += (int)t11 * (fyi * fxi);
                  const int qv = S >> 10, rem = S & 1023;
                  v = u2f_exact((uint32_t)(qv + ((rem + (qv & 1)) > 512)));
                } else {
                  const float fx = frac32_exact(fxi), fy = frac32_exact(fyi);
                  const float gx = 1.0f - fx, gy = 1.0f - fy;
                  float r = __fadd_rn(__fmul_rn((float)t00, __fmul_rn(gy, gx)), __fmul_rn((float)t01, __fmul_rn(gy, fx)));
                  r = __fadd_rn(r, __fmul_rn((float)t10, __fmul_rn(fy, gx)));
                  r = __fadd_rn(r, __fmul_rn((float)t11, __fmul_rn(fy, fx)));
                  r = __fadd_rn(__fadd_rn(r, 12582912.0f), -12582912.0f);
                  v = fminf(r, 65535.0f);
                }
              } else {
                v = warp_px_slow(fr, W, H, X, Y, a.interp);
              }
            }
            const float sol = __fadd_rn(0.0f, __fmul_rn(val, v));
            tile[u0 + j][threadIdx.x] = sol;
            q += (double)__fmul_rn(sol, sol);
            s += (double)sol;
          }
        } else {
          for (int j = 0; j < ns; ++j) {
            float sol = __int_as_float(0x7fc00000);     // skipped node
            if (code <= -2) sol = __fadd_rn(0.0f, __fmul_rn(val, __ldg(cam.pv + (size_t)(-2 - code) * a.bstride + b + j)));
            tile[u0 + j][threadIdx.x] = sol;
            q += (double)__fmul_rn(sol, sol);
            s += (double)sol;
          }
        }
      }
    }
    __syncthreads();
    if (lane < nb) {
#pragma unroll 4
      for (int j = 0; j < 32; ++j) {
        float* rp = rowp[w * 32 + j];
        if (rp != nullptr) rp[b0 + lane] = tile[lane][w * 32 + j];
      }
    }
    __syncthreads();
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
