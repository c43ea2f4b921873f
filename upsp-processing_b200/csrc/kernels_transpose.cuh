// kernels_transpose.cuh -- K6: frame-major -> node-major transpose, fused with the
// all-to-all (stores go straight into the destination rank's [N_s x F] buffer).
#pragma once
#include "common.cuh"

namespace upsp {

// Reference: local_transpose cpp/exec/psp_process.cpp:647-689 (100x100 CPU tiles) +
// global_transpose :707-771 (Isend of [N_s x F_r] blocks, Recv, reassembly copy into
// dst[node][start_col + f] :755-765).  Here one kernel reads this rank's [F_r x N] slice
// once and writes every element to its final position dst_s[n - n0_s][f0_r + f]; when
// dst_s is a peer-mapped pointer the store crosses NVLink (no bounce buffer, no
// reassembly pass).  64x64 tiles staged through shared memory so both the global read
// (along nodes) and the global write (along frames) are 128-bit and fully coalesced.
struct XposeArgs {
  const float* src;  // [rows][cols] = [F_r][N]
  int rows, cols;
  int n_ranks;
  int f_total;       // row length of the destination (number_frames)
  int col0;          // rank_start_frame[this rank]
  float* dst[UPSP_MAX_RANKS];          // base of rank s's node-major buffer
  int node_start[UPSP_MAX_RANKS + 1];  // rank_start_node[], [n_ranks] = N
};

constexpr int XT = 64;  // tile edge

__global__ void __launch_bounds__(256)
k_transpose_a2a(const XposeArgs a) {
  __shared__ float tile[XT][XT + 1];
  const int n0 = blockIdx.x * XT;  // node (src column) origin
  const int f0 = blockIdx.y * XT;  // frame (src row) origin
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 floats each
  const bool vec_in = (a.cols & 3) == 0;
  // load: rows f0+ty+16*j, cols n0+4*tx..+3
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int f = f0 + ty + 16 * j, n = n0 + 4 * tx;
    if (f < a.rows) {
      const float* p = a.src + (size_t)f * a.cols + n;
      if (vec_in && n + 3 < a.cols) {
        float4 v = ld_stream_f4(p);
        tile[ty + 16 * j][4 * tx + 0] = v.x;
        tile[ty + 16 * j][4 * tx + 1] = v.y;
        tile[ty + 16 * j][4 * tx + 2] = v.z;
        tile[ty + 16 * j][4 * tx + 3] = v.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (n + k < a.cols) tile[ty + 16 * j][4 * tx + k] = p[k];
      }
    }
  }
  __syncthreads();
  // store: node n0+ty+16*j, frames f0+4*tx..+3
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + ty + 16 * j, f = f0 + 4 * tx;
    if (n >= a.cols || f >= a.rows) continue;
    int s = 0;
    while (s + 1 < a.n_ranks && n >= a.node_start[s + 1]) ++s;
    float* q = a.dst[s] + (size_t)(n - a.node_start[s]) * a.f_total + a.col0 + f;
    float4 v = make_float4(tile[4 * tx + 0][ty + 16 * j], tile[4 * tx + 1][ty + 16 * j],
                           tile[4 * tx + 2][ty + 16 * j], tile[4 * tx + 3][ty + 16 * j]);
    if (f + 3 < a.rows && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
      st_stream_f4(q, v);
    } else {
      q[0] = v.x;
      if (f + 1 < a.rows) q[1] = v.y;
      if (f + 2 < a.rows) q[2] = v.z;
      if (f + 3 < a.rows) q[3] = v.w;
    }
  }
}

// ---- staged exchange, shipped by the SMs: rows of the other ranks' nodes, which the projection kernel left in the
// local staging block [N][stride] (column = frame inside the batch), go to columns [col0, col0 + nb) of their owners'
// node-major buffers [N_s][F].  A batch makes nb * 4 bytes (1 KB at 256 frames) of every row contiguous at the
// destination, against 64 / 128 bytes when the projection kernel stores there itself; the kernel is small (a few warps
// per SM) and runs beside the projection of the next batch, so neither the projection nor the copy engines (290 GB/s
// per GPU, measured) sit on the NVLink path.  Reference: the MPI_Alltoallv of global_transpose,
// cpp/exec/psp_process.cpp:707-771.  Peers are interleaved row by row so that every link is busy all the time.
struct ShipArgs {
  const float* stage;
  int stride, nb, n_peers, max_rows;
  float* dst[UPSP_MAX_RANKS];        // owner's buffer, already offset to column col0 of its row 0
  int row0[UPSP_MAX_RANKS];          // first node of the owner in the staging block
  int rows[UPSP_MAX_RANKS];
  size_t dst_stride;                 // F (floats)
};

__global__ void __launch_bounds__(128)
k_ship_rows(const ShipArgs a) {
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const bool vec = (a.nb & 3) == 0 && (a.stride & 3) == 0;
  for (int r = w; r < a.max_rows; r += warps) {
    for (int p = 0; p < a.n_peers; ++p) {
      if (r >= a.rows[p]) continue;
      const float* src = a.stage + (size_t)(a.row0[p] + r) * a.stride;
      float* dst = a.dst[p] + (size_t)r * a.dst_stride;
      if (vec && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        for (int i = lane * 4; i < a.nb; i += 128) st_stream_f4(dst + i, ld_stream_f4(src + i));
      } else {
        for (int i = lane; i < a.nb; i += 32) dst[i] = src[i];
      }
    }
  }
}

}  // namespace upsp
