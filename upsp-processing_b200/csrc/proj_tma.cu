// proj_tma.cu -- translation unit of the TMA-staged projection kernels (kernels_project_tma.cuh).
#include "proj_tma.h"

#include <cstdlib>
#include <cstring>

#include "kernels_project_tma.cuh"

namespace upsp {

TmaGeom tma_geom() {
  TmaGeom g;
  g.nodes_per_block = TMA_NB;
  g.strip_rows = TMA_TH;
  g.tile_cols = TMA_TW;
  g.box_rows = TMA_BH;
  g.box_px16 = TMA_BW16;
  g.box_words12 = TMA_BWB12 / 4;
  g.group_frames = TMA_G;
  return g;
}

template <int SRC, int CH, bool VAL1, int NG = TMA_NG, int LA = TMA_LA, int MINB = 6, bool IT16 = false>
static cudaError_t launch1(const CUtensorMap& map_group, const CUtensorMap& map_single, const FusedArgs& a, const TmaExtra& ex,
                           int nblocks, cudaStream_t st) {
  constexpr int smem = TmaSmem<SRC, CH, NG>::total;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k_project_tma<SRC, CH, VAL1, NG, LA, MINB, IT16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const unsigned ny = ex.split_frames > 0 ? (unsigned)((a.nframes + ex.split_frames - 1) / ex.split_frames) : 1u;
  k_project_tma<SRC, CH, VAL1, NG, LA, MINB, IT16><<<dim3((unsigned)nblocks, ny), TMA_NB, smem, st>>>(map_group, map_single, a, ex);
  return cudaGetLastError();
}

cudaError_t launch_project_tma(int src, bool seg128, bool val1, bool rows16, const CUtensorMap& map_group,
                               const CUtensorMap& map_single, const FusedArgs& a, const TmaExtra& ex, int nblocks, cudaStream_t st) {
  if (nblocks <= 0) return cudaSuccess;
  if (rows16) {      // 16-bit node-major rows (unit projection values only)
    if (!val1) return cudaErrorInvalidValue;
    if (src == 0) return seg128 ? launch1<0, 32, true, 3, 2, 6, true>(map_group, map_single, a, ex, nblocks, st)
                                : launch1<0, 16, true, TMA_NG, TMA_LA, 6, true>(map_group, map_single, a, ex, nblocks, st);
    return seg128 ? launch1<1, 32, true, 3, 2, 6, true>(map_group, map_single, a, ex, nblocks, st)
                  : launch1<1, 16, true, TMA_NG, TMA_LA, 6, true>(map_group, map_single, a, ex, nblocks, st);
  }
  // UPSP_TMA_RING = "NG.LA.MINB" picks another ring depth / look-ahead / occupancy target of the hot instantiation
  // (packed source, 64-byte segments, unit values): tuning knob
  static const char* ring = getenv("UPSP_TMA_RING");
  if (ring && src == 1 && !seg128 && val1) {
#define UPSP_RING(NG, LA, MB) \
    if (!strcmp(ring, #NG "." #LA "." #MB)) return launch1<1, 16, true, NG, LA, MB>(map_group, map_single, a, ex, nblocks, st)
    UPSP_RING(3, 2, 7);
    UPSP_RING(3, 1, 7);
    UPSP_RING(3, 2, 6);
    UPSP_RING(4, 3, 6);
    UPSP_RING(5, 3, 5);
    UPSP_RING(6, 4, 5);
    UPSP_RING(4, 2, 5);
    UPSP_RING(2, 1, 8);
#undef UPSP_RING
  }
  // 128-byte-segment variants (C = 32): 3-group ring, see TmaSmem
#define UPSP_TMA_CASE(S, C)                                                                              \
  return val1 ? launch1<S, C, true, (C == 32 ? 3 : TMA_NG), TMA_LA>(map_group, map_single, a, ex, nblocks, st)   \
              : launch1<S, C, false, (C == 32 ? 3 : TMA_NG), TMA_LA>(map_group, map_single, a, ex, nblocks, st)
  if (src == 0) {
    if (seg128) UPSP_TMA_CASE(0, 32);
    UPSP_TMA_CASE(0, 16);
  }
  if (seg128) UPSP_TMA_CASE(1, 32);
  UPSP_TMA_CASE(1, 16);
#undef UPSP_TMA_CASE
}

cudaError_t launch_hot_scan12(const uint8_t* in, size_t in_stride, size_t npix, int nframes, int thresh, int* hot_cnt,
                              int* hot_pos, int* done, int rows, int cols, void* fixes, int grid, int threads, cudaStream_t st) {
  static const int env_threads = getenv("UPSP_SCAN_THREADS") ? atoi(getenv("UPSP_SCAN_THREADS")) : 0;   // tuning knob
  const int scan_threads = env_threads > 0 ? env_threads : threads;
  k_hot_scan12<<<grid, scan_threads, 0, st>>>(in, in_stride, npix, nframes, thresh, hot_cnt, hot_pos, done, rows, cols,
                                     reinterpret_cast<HotFix*>(fixes));
  return cudaGetLastError();
}

size_t hot_fix_bytes() { return sizeof(HotFix); }

cudaError_t tma_set_coef(int set, const double2* dev_coef, int n, cudaStream_t st) {
  return cudaMemcpyToSymbolAsync(c_tma_coef, dev_coef, (size_t)n * sizeof(double2), (size_t)set * TMA_MAXB * sizeof(double2),
                                 cudaMemcpyDeviceToDevice, st);
}
int tma_max_batch() { return TMA_MAXB; }
int tma_stage_frames() { return TMA_S; }

}  // namespace upsp
