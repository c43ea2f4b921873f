// kernels_project.cuh -- K5: batched-frame CSR projection (pixel -> node gather) with the
// camera blend, NaN fill, overlap remap, sum / sum-of-squares and row store fused in.
#pragma once
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
  return code >= 0 ? u2f_exact(__ldg(c.frames + (size_t)b * c.npix + code))
                   : __ldg(c.pv + (size_t)(-2 - code) * bstride + b);
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
  const float* pv;         // [slots][bstride] patch values
  const int* code;         // [N]
  const float* val;        // [N]
};
struct FusedArgs {
  int n_cams, n_nodes, nframes, bstride, interp, skip_frame;
  FusedCam cam[UPSP_MAX_CAMS];
  double* sum;
  double* sumsq;
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

template <int NC, bool REG>
__global__ void __launch_bounds__(256)
k_project_fused(const FusedArgs a) {
  __shared__ float tile[32][257];
  const int n = blockIdx.x * 256 + threadIdx.x;
  const bool live = n < a.n_nodes;
  int code[NC], tx[NC], ty[NC];
  float val[NC];
  bool skipped = true;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    code[c] = live ? __ldg(a.cam[c].code + n) : -1;
    val[c] = live ? __ldg(a.cam[c].val + n) : 0.0f;
    skipped = skipped && (code[c] == -1);
    const int W = a.cam[c].W;
    const int px = code[c] >= 0 ? code[c] % W : 0, py = code[c] >= 0 ? code[c] / W : 0;
    tx[c] = px;              // int2 index of (adelta,bdelta)[px] inside a frame's table
    ty[c] = W + py;          // int2 index of (X0,Y0)[py]
  }
  double s = 0.0, q = 0.0;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b0 = 0; b0 < a.nframes; b0 += 32) {
    const int nb = min(32, a.nframes - b0);
    if (live) {
#pragma unroll 4
      for (int u = 0; u < nb; ++u) {
        const int b = b0 + u;
        float sol = 0.0f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float cs = 0.0f;
          const FusedCam& cam = a.cam[c];
          if (code[c] >= 0) {
            float v;
            const unsigned fbase = (unsigned)b * (unsigned)cam.npix;   // batch-local pixel index (< 2^31)
            if (REG && b != a.skip_frame) {
              const int2* tb = reinterpret_cast<const int2*>(cam.tab) + (unsigned)b * (unsigned)(cam.W + cam.H);
              const int2 xa = __ldg(tb + tx[c]);
              const int2 ya = __ldg(tb + ty[c]);
              const int X = ya.x + xa.x, Y = ya.y + xa.y;
              const int Xs = X >> 5, Ys = Y >> 5;
              const int sx = X >> 10, sy = Y >> 10;
              if (a.interp == 1 && (unsigned)sx < (unsigned)(cam.W - 1) && (unsigned)sy < (unsigned)(cam.H - 1)) {
                const uint16_t* p = cam.frames + (fbase + (unsigned)(sy * cam.W + sx));
                const uint16_t* p2 = p + cam.W;
                const float t00 = (float)__ldg(p), t01 = (float)__ldg(p + 1);
                const float t10 = (float)__ldg(p2), t11 = (float)__ldg(p2 + 1);
                const float fx = frac32_exact(Xs & 31), fy = frac32_exact(Ys & 31);
                const float gx = 1.0f - fx, gy = 1.0f - fy;
                v = __fadd_rn(__fmul_rn(t00, __fmul_rn(gy, gx)), __fmul_rn(t01, __fmul_rn(gy, fx)));
                v = __fadd_rn(v, __fmul_rn(t10, __fmul_rn(fy, gx)));
                v = __fadd_rn(v, __fmul_rn(t11, __fmul_rn(fy, fx)));
                v = __fadd_rn(__fadd_rn(v, 12582912.0f), -12582912.0f);   // rint (half to even)
                v = fminf(v, 65535.0f);                                    // v >= 0 by construction
              } else {
                v = warp_px_slow(cam.frames + fbase, cam.W, cam.H, X, Y, a.interp);
              }
            } else {
              v = (float)__ldg(cam.frames + (fbase + (unsigned)code[c]));
            }
            cs = __fadd_rn(0.0f, __fmul_rn(val[c], v));
          } else if (code[c] <= -2) {
            cs = __fadd_rn(0.0f, __fmul_rn(val[c], __ldg(cam.pv + (size_t)(-2 - code[c]) * a.bstride + b)));
          }
          sol = (c == 0) ? cs : __fadd_rn(sol, cs);
        }
        if (skipped) sol = __int_as_float(0x7fc00000);
        tile[u][threadIdx.x] = sol;
        q += (double)__fmul_rn(sol, sol);
        s += (double)sol;
      }
    }
    __syncthreads();
    // warp w writes nodes [w*32, w*32+32) of the block: lane = frame -> 128-byte row segments
    {
      const int n0 = blockIdx.x * 256 + w * 32;
      const int n1 = min(n0 + 32, a.n_nodes);
      int r = 0;
      while (r + 1 < a.n_ranks && n0 >= a.node_start[r + 1]) ++r;
      const size_t col = (size_t)a.col0 + b0 + lane;
      float* p = a.dst[r] + (size_t)(n0 - a.node_start[r]) * a.f_total + col;
      int next = a.node_start[r + 1];
      for (int nn = n0; nn < n1; ++nn) {
        if (nn >= next) {   // the warp's node run crosses into the next rank's slice
          do {
            ++r;
            next = a.node_start[r + 1];
          } while (nn >= next);
          p = a.dst[r] + (size_t)(nn - a.node_start[r]) * a.f_total + col;
        }
        if (lane < nb) *p = tile[lane][nn - blockIdx.x * 256];
        p += a.f_total;
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
