// tma_bench.cu -- TMA box-load throughput per SM as a function of the box shape (u16 3-D map [F][H][W]).
// Every block keeps DEPTH box loads in flight (ring of mbarriers) and does nothing else.
// usage: tma_bench BW BH BF blocks_per_sm depth iters [l2promo 0|1|2|3]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_bench(const __grid_constant__ CUtensorMap tmap, int box_bytes, int depth, int iters, int W, int H, int F,
                        int BW, int BH, int BF, unsigned long long* sink, int issuers, int lane_mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  unsigned char* ring = smem + 1024;
  const int slot_bytes = (box_bytes + 127) & ~127;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth * issuers; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + i)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int me = -1;
  if (lane_mode) { if ((int)threadIdx.x < issuers) me = threadIdx.x; }
  else if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < issuers) me = threadIdx.x >> 5;
  if (me < 0) return;
  bar += me * depth;
  ring += (size_t)me * depth * slot_bytes;
  // walk the image like the projection does: block b owns a tile, frames advance
  const int tiles_x = (W - BW) / 80 + 1, tiles_y = (H - BH) / 4 + 1;
  const int t = blockIdx.x % (tiles_x * tiles_y);
  const int x0 = (t % tiles_x) * 80, y0 = (t / tiles_x) * 4;
  unsigned long long acc = 0;
  for (int i = 0; i < iters + depth; ++i) {
    const int s = i % depth;
    if (i >= depth) {
      const unsigned par = ((i / depth) - 1) & 1;
      asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
                       smem_u32(bar + s)),
                   "r"(par)
                   : "memory");
      acc += ring[s * slot_bytes];
    }
    if (i < iters) {
      const int f = (i * BF) % (F - BF + 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + s)), "r"(box_bytes) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
              smem_u32(ring + s * slot_bytes)),
          "l"(&tmap), "r"(smem_u32(bar + s)), "r"(x0), "r"(y0), "r"(f)
          : "memory");
    }
  }
  if (acc == 0x123456789ull) *sink = acc;
}

int main(int argc, char** argv) {
  if (argc < 7) return 2;
  const int BW = atoi(argv[1]), BH = atoi(argv[2]), BF = atoi(argv[3]), bps = atoi(argv[4]), depth = atoi(argv[5]), iters = atoi(argv[6]);
  const int promo = argc > 7 ? atoi(argv[7]) : 2;
  const int issuers = argc > 8 ? atoi(argv[8]) : 1;
  const int lane_mode = argc > 9 ? atoi(argv[9]) : 0;
  const int W = 1024, H = 1024, F = 256;
  uint16_t* d;
  cudaMalloc(&d, (size_t)W * H * F * 2);
  cudaMemset(d, 1, (size_t)W * H * F * 2);
  unsigned long long* sink;
  cudaMalloc(&sink, 8);
  CUresult (*enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  enc = reinterpret_cast<decltype(enc)>(fn);
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 2, (cuuint64_t)W * H * 2};
  const cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BF};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode -> %d\n", (int)r);
    return 1;
  }
  const int box_bytes = BW * BH * BF * 2;
  const int smem = 1024 + issuers * depth * ((box_bytes + 127) & ~127);
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_bench<<<nsm * bps, lane_mode ? 32 : 32 * issuers, smem>>>(m, box_bytes, depth, iters, W, H, F, BW, BH, BF, sink, issuers, lane_mode);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("box %dx%dx%d: %s\n", BW, BH, BF, cudaGetErrorString(e));
      return 1;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 2) {
      const double boxes_per_sm = (double)bps * iters * issuers;
      const double cyc = ms * 1e-3 * 1.965e9;
      printf("box %3dx%2dx%d (%5d B, %2d rows) bps %d x %d issuers (%s) depth %2d promo %d: %.3f ms, %.1f cyc/box/SM, %.1f cyc/row, %.1f B/cyc/SM, %.2f TB/s chip\n", BW, BH,
             BF, box_bytes, BH * BF, bps, issuers, lane_mode ? "lanes" : "warps", depth, promo, ms, cyc / boxes_per_sm, cyc / boxes_per_sm / (BH * BF), box_bytes * boxes_per_sm / cyc,
             box_bytes * boxes_per_sm * nsm / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
