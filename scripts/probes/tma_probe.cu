// tma_probe.cu -- which (coordinate, box, alignment) combinations the TMA unit accepts for a 3-D u16 map.
// usage: tma_probe W H F BW BH x y f [smem_off]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int f, int bytes, int smem_off, uint16_t* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  unsigned char* dst = smem + 1024 + smem_off;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(&tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(f)
        : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(0)
      : "memory");
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(dst)[i];
}

int main(int argc, char** argv) {
  if (argc < 9) return 2;
  const int W = atoi(argv[1]), H = atoi(argv[2]), F = atoi(argv[3]), BW = atoi(argv[4]), BH = atoi(argv[5]);
  const int x = atoi(argv[6]), y = atoi(argv[7]), f = atoi(argv[8]);
  const int smem_off = argc > 9 ? atoi(argv[9]) : 0;
  std::vector<uint16_t> h((size_t)W * H * F);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i % 4001);
  uint16_t *d, *dout;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&dout, BW * BH * 2);
  cudaMemset(dout, 0xff, BW * BH * 2);
  CUresult (*enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  enc = reinterpret_cast<decltype(enc)>(fn);
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 2, (cuuint64_t)W * H * 2};
  const cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode -> %d\n", (int)r);
    return 1;
  }
  k_probe<<<1, 64, 1024 + smem_off + BW * BH * 2 + 256>>>(m, x, y, f, BW * BH * 2, smem_off, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("W%d H%d F%d box %dx%d at (%d,%d,%d) smem+%d: %s\n", W, H, F, BW, BH, x, y, f, smem_off, cudaGetErrorString(e));
    return 1;
  }
  std::vector<uint16_t> o((size_t)BW * BH);
  cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < BH; ++r2)
    for (int c = 0; c < BW; ++c) {
      const int xx = x + c, yy = y + r2;
      const uint16_t want = (xx >= 0 && xx < W && yy >= 0 && yy < H && f >= 0 && f < F) ? h[((size_t)f * H + yy) * W + xx] : 0;
      if (o[(size_t)r2 * BW + c] != want) ++bad;
    }
  printf("W%d H%d F%d box %dx%d at (%d,%d,%d) smem+%d: ok, %d mismatches\n", W, H, F, BW, BH, x, y, f, smem_off, bad);
  return 0;
}
