// proj_tma.h -- host interface of the TMA-staged projection kernels (proj_tma.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "project_args.cuh"

namespace upsp {

// geometry of the staged boxes (the host builds the block partition and the tensor maps from these)
struct TmaGeom {
  int nodes_per_block;   // TMA_NB
  int strip_rows;        // TMA_TH
  int tile_cols;         // TMA_TW: widest column span of a block's node pixels
  int box_rows;          // TMA_BH
  int box_px16;          // SRC 0: box width in u16 pixels
  int box_words12;       // SRC 1: box width in 32-bit words of packed bytes
  int group_frames;      // TMA_G: frames fetched by one box when their boxes coincide
};
TmaGeom tma_geom();

struct TmaBlock {        // nodes [node0, node0+count) of the processing order, all with a plain pixel
  int node0, count;
  int xmin, xmax;        // column range of the block's node pixels (xmax - xmin <= tile_cols)
  int ymin, ymax;        // row range (ymax - ymin < strip_rows)
  int pad0, pad1;
};

struct HotFix;           // pixel_ops.cuh

struct TmaExtra {
  const TmaBlock* blk;
  const uint8_t* packed;        // SRC 1: packed frames of the batch (slow path), stride frame_bytes
  size_t frame_bytes;
  int frame0;                   // slot of the batch's first frame in the tensor map (third TMA coordinate)
  const HotFix* hot;            // SRC 1: [batch] fix lists (nullptr: hot-pixel fix off)
  int split_frames;             // > 0: grid.y slices of this many frames each (integer statistics only), 0: one block per node tile
  int coef_set;                 // which of the two constant-memory coefficient tables holds this batch (tma_set_coef)
};

// src 0: decoded u16 frames, 1: packed 12-bit frames.  seg128: 128-byte row segments (peer stores).
// map_group: box of group_frames frames; map_single: box of one frame (same tensor, same rows x columns).
// rows16: node-major rows stored as 16-bit integers (val1 only; dst[] are then 2-byte element buffers).
cudaError_t launch_project_tma(int src, bool seg128, bool val1, bool rows16, const CUtensorMap& map_group,
                               const CUtensorMap& map_single, const FusedArgs& a, const TmaExtra& ex, int nblocks, cudaStream_t st);
// hot-pixel scan of packed 12-bit frames -> fix lists (read only)
cudaError_t launch_hot_scan12(const uint8_t* in, size_t in_stride, size_t npix, int nframes, int thresh, int* hot_cnt,
                              int* hot_pos, int* done, int rows, int cols, void* fixes, int grid, int threads, cudaStream_t st);
size_t hot_fix_bytes();
// (M0, M3) * 1024 of `n` frames (device array of double2, identity for the unregistered frame) -> constant table `set`
cudaError_t tma_set_coef(int set, const double2* dev_coef, int n, cudaStream_t st);
int tma_max_batch();
int tma_stage_frames();

}  // namespace upsp
