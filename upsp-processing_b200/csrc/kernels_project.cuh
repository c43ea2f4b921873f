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
