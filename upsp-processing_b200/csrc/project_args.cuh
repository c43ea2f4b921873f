// project_args.cuh -- argument blocks of the projection kernels (shared by the translation units
// that hold them: upsp_gpu.cu and proj_tma.cu).
#pragma once
#include "common.cuh"

namespace upsp {

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
  FusedCam cam[UPSP_MAX_CAMS];
  double* sum;
  double* sumsq;
  const int* perm;                     // [N] processing order: nodes sorted by pixel index (raster),
                                       // so a warp gathers from one or two image rows
  int n_ranks, f_total, col0;          // col0 = global frame index of the batch's first frame
  float* dst[UPSP_MAX_RANKS];          // node-major [N_s][F] buffer of every rank
  int node_start[UPSP_MAX_RANKS + 1];
  // staged exchange (n_ranks > 1, pipelined): rows of nodes owned by the ranks in `stage_mask` are
  // written to a local [N][stage_stride] staging block (column = frame inside the batch) and shipped
  // to their owners by copy engines afterwards; the other ranks' rows go straight into the
  // peer-mapped buffers.  Copy engines and SM stores then drive NVLink side by side.
  float* stage;
  int stage_stride, rank;
  unsigned stage_mask;                 // bit r set: rank r's rows go through the staging block
  // 16-bit row mode: the few nodes whose values are not integers (patched pixels) or not numbers (unseen nodes) keep
  // float rows in a side buffer of their owner; row_index[n] = row of node n in that buffer (dst[] then point at the
  // side buffers).  nullptr: row = n - node_start[owner].
  const int* row_index;
  // 16-bit rows, batch-blocked destination (multi-rank): rank s holds [block][N_s][blk_len] values, block = source rank *
  // blocks per rank + local batch; this launch writes block blk_index from frame blk_j0 of the block on.  blk_len = 0:
  // node-major rows [N_s][f_total].
  int blk_len, blk_index, blk_j0;
};

// where a block writes node n's row segment of this batch (frame b of the batch at [b])
__device__ __forceinline__ float* fused_row_ptr(const FusedArgs& a, int n) {
  int r = 0;
  while (r + 1 < a.n_ranks && n >= a.node_start[r + 1]) ++r;
  if (a.stage != nullptr && ((a.stage_mask >> r) & 1u)) return a.stage + (size_t)n * a.stage_stride;
  const int row = a.row_index != nullptr ? __ldg(a.row_index + n) : n - a.node_start[r];
  return a.dst[r] + (size_t)row * a.f_total + a.col0;
}

}  // namespace upsp
