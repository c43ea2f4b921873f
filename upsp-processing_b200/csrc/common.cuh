// common.cuh -- shared device helpers for libupsp_gpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef UPSP_MAX_CAMS
#define UPSP_MAX_CAMS 8
#endif
#define UPSP_MAX_RANKS 16
#define UPSP_HOT_STORE 8      // positions remembered per frame (max_hot = 5)
#define UPSP_HOT_THRESH 4064  // cpp/include/utils/cv_extras.h:154-155
#define UPSP_HOT_MIN_CHANGE 512
#define UPSP_HOT_MAX 5
#define UPSP_MAX_COEF 9       // detrend degree <= 8

namespace upsp {

// streaming (touch-once) 128-bit global accesses: keep them out of L1
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  uint4 r = ld_stream_u4(p);
  return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z),
                     __uint_as_float(r.w));
}
__device__ __forceinline__ void st_stream_f4(void* p, const float4& v) {
  st_stream_u4(p, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z),
                             __float_as_uint(v.w)));
}

// Conversions without the (quarter-rate) I2F / F2I pipe: exact for the stated ranges.
__device__ __forceinline__ float u2f_exact(uint32_t v) {        // v < 2^23
  return __fadd_rn(__uint_as_float(0x4B000000u | v), -8388608.0f);
}
__device__ __forceinline__ float frac32_exact(uint32_t v) {     // v in [0,32) -> v/32
  return __fadd_rn(__uint_as_float(0x48800000u | v), -262144.0f);
}
__device__ __forceinline__ int f2i_rn_small(float v) {           // |v| < 2^22, round half to even
  return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace upsp
