// kernels_project_tma.cuh -- K3a + K5 + K6 for the hot configuration (one camera, bilinear
// registration, 12-bit containers), with the source pixels staged in shared memory by TMA.
//
// Reference: cv::warpAffine at the projected pixels (cpp/lib/registration.cpp:69-72), project_frame
// (cpp/lib/projection.ipp:884-908), NaN / sums / row store (cpp/exec/psp_process.cpp:1814-1842),
// local_transpose + global_transpose (:647-771).  Same arithmetic, same bits as k_project_fused4.
//
// Why: k_project_fused4 gathers its four taps per node-frame straight from global memory and is
// bound by the latency of those loads (ncu r01: long-scoreboard 5.4 stall cycles per issue, 0.34 of
// the HBM peak at 1.05x the algorithmic traffic).  Here a block owns up to 128 nodes whose pixels
// lie in one TH-row strip segment of the image; for every frame ONE elected thread asks the TMA unit
// for the [BH rows x BW px] box those nodes can touch (cp.async.bulk.tensor, 3-D map over
// [frame][row][px], completion on an mbarrier), NG groups of 4 frames ahead of the consumers.  The
// taps then are shared-memory loads (29 cycles, no long scoreboard), out-of-image taps are the TMA
// unit's zero fill (= cv::BORDER_CONSTANT 0, so the border needs no slow path), and the unregistered
// frame (global frame 0, psp_process.cpp:1777) is the identity map through the same code.
//
// SRC = 0: the box is cut from the decoded u16 frames (k_unpack12_scan* output);
// SRC = 1: the box is cut from the PACKED 12-bit frames as pushed (3 bytes = 2 px, MSB first,
//          cpp/lib/PSPVideo.cpp:134-150) and the taps are unpacked on the fly: the decoded frame is
//          never written (saves 2P write + 2P read per frame and the decode pass).  The <= 5 hot-pixel
//          fixes of a frame (cpp/utils/cv_extras.cpp:230-272) come as a (position, value) list from
//          k_hot_scan12 and are patched into the staged box before it is consumed.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "pixel_ops.cuh"
#include "project_args.cuh"
#include "proj_tma.h"

namespace upsp {

constexpr int TMA_NB = 128;          // nodes (= threads) per block
constexpr int TMA_TH = 4;            // image rows a block's node pixels may span
constexpr int TMA_G = 4;             // frames per group (one mbarrier phase, ONE box when the frames' boxes coincide)
constexpr int TMA_BH = TMA_TH + 4;   // rows of the staged box (bilinear +1, rounding of the warp +1, +2: frames of a group move apart)
constexpr int TMA_BW16 = 96;         // SRC 0: box width in pixels (192 B rows)
constexpr int TMA_BWB12 = 176;       // SRC 1: box width in bytes (117 px)
constexpr int TMA_TW = 80;           // widest column span of a block's node pixels (both sources)
constexpr int TMA_NG = 4;            // groups in the ring
constexpr int TMA_LA = 2;            // groups the producer runs ahead of the consumers (<= NG - 1; NG - LA groups of slack between warps)
constexpr int TMA_S = 64;            // frames per table stage
constexpr int TMA_SLOT16 = TMA_BH * TMA_BW16 * 2;   // bytes per staged frame = the box plane, so that a box of G frames fills G slots
constexpr int TMA_SLOT12 = TMA_BH * TMA_BWB12;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // try_wait with a suspend-time hint: a waiting warp sleeps in the barrier unit instead of spinning through the issue slots
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(20000u)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Per-frame column coefficients (M0, M3) * 1024 of the batch (identity for the unregistered global frame 0), in
// CONSTANT memory: every lane of a warp reads the same frame's pair, and as shared-memory loads those broadcasts
// cost two wavefronts each of the load/store data pipe this kernel is bound by (ncu r2i: 76 % of its peak, a tenth
// of it these loads).  Two sets: the copy for batch i+1 (front-end stream) runs under the kernel of batch i.
constexpr int TMA_MAXB = 1024;       // largest batch the table holds (host-checked)
__constant__ double2 c_tma_coef[2][TMA_MAXB];

// cv::warpAffine at one pixel from a packed frame, any coordinates (the box of a frame did not fit)
static __device__ __noinline__ float warp_px_slow12(const uint8_t* __restrict__ fr, const HotFix* __restrict__ h, int W, int H,
                                             int X, int Y) {
  const int Xs = X >> 5, Ys = Y >> 5;
  const int sx = Xs >> 5, sy = Ys >> 5;
  const int fxi = Xs & 31, fyi = Ys & 31;
  uint32_t t[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xx = sx + (k & 1), yy = sy + (k >> 1);
    t[k] = ((unsigned)xx < (unsigned)W && (unsigned)yy < (unsigned)H) ? px_packed12_fixed(fr, (unsigned)(yy * W + xx), h) : 0u;
  }
  const unsigned S = (t[0] * (32 - fxi) + t[1] * fxi) * (32 - fyi) + (t[2] * (32 - fxi) + t[3] * fxi) * fyi;
  return __fadd_rn(__fmaf_rn(__uint_as_float(S + 0x4B000000u), 0.0009765625f, 12574720.0f), -12582912.0f);
}

// cv::warpAffine at fixed-point coordinates (X, Y) RELATIVE TO THE STAGED BOX (INTER_BITS 5, AB_BITS 10), taps from
// shared memory: returns the float 1.5 * 2^23 + round_half_even(S / 1024), S = sum t_ij w_ij (integer, exact)
template <int SRC>
__device__ __forceinline__ float tma_px(const unsigned char* __restrict__ sl, int X, int Y) {
  const unsigned fxi = ((unsigned)X >> 5) & 31u, fyi = ((unsigned)Y >> 5) & 31u;
  const unsigned gy = 32u - fyi;
  unsigned top, bot;
  if (SRC == 0) {
    const unsigned short* p = reinterpret_cast<const unsigned short*>(sl) + ((Y >> 10) * TMA_BW16 + (X >> 10));
    const unsigned t00 = p[0], t01 = p[1], t10 = p[TMA_BW16], t11 = p[TMA_BW16 + 1];
    const unsigned gx = 32u - fxi;
    top = t00 * gx + t01 * fxi;
    bot = t10 * gx + t11 * fxi;
  } else {
    // pixels dx, dx+1 of a packed row: 24 bits from bit 12*dx, MSB first
    const unsigned dx = (unsigned)(X >> 10);
    const unsigned o = dx + (dx >> 1);                        // byte offset floor(1.5 dx)
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(sl + (Y >> 10) * TMA_BWB12 + (o & ~3u));
    const unsigned sel = 0x0123u + 0x1111u * (o & 3u);        // bytes o..o+3, big-endian
    const unsigned sh = (dx & 1u) * 4u;
    const unsigned r0 = __byte_perm(pw[0], pw[1], sel) << sh;
    const unsigned r1 = __byte_perm(pw[TMA_BWB12 / 4], pw[TMA_BWB12 / 4 + 1], sel) << sh;
    // r = px0 << 20 | px1 << 8 | junk; px0 gx + px1 fxi = (r >> 8) fxi + px0 (gx - 4096 fxi)
    const int cx = 32 - 4097 * (int)fxi;
    top = (unsigned)((int)((r0 >> 8) * fxi) + (int)(r0 >> 20) * cx);
    bot = (unsigned)((int)((r1 >> 8) * fxi) + (int)(r1 >> 20) * cx);
  }
  const unsigned sm = top * gy + 0x4B000000u + bot * fyi;      // bits of the float 2^23 + S
  return __fmaf_rn(__uint_as_float(sm), 0.0009765625f, 12574720.0f);
}

// shared-memory carve-up (dynamic, 128-byte aligned base)
static_assert(TMA_SLOT16 % 128 == 0 && TMA_SLOT12 % 128 == 0, "staged boxes must stay 128-byte aligned");
static_assert(TMA_G == 4 && TMA_S % 32 == 0, "a group's frames are four neighbouring lanes of the stage set-up");
static_assert(TMA_BWB12 % 16 == 0 && (TMA_BW16 * 2) % 16 == 0, "box rows are multiples of 16 bytes");
template <int SRC, int CH, int NG = TMA_NG>
struct TmaSmem {
  static constexpr int SLOT = SRC ? TMA_SLOT12 : TMA_SLOT16;
  // tile row = CH floats (or 2 CH 16-bit values).  CH = 16: + 16 bytes of padding (80-byte rows).  CH = 32 (128-byte row
  // segments for peer stores): no padding, the eight 16-byte chunks of a row are XOR-swizzled with the row index
  // instead, which keeps the block at 37 KB with a 3-group ring: six blocks per SM like the 64-byte variant (with the
  // padded 18 KB tile + 4-group ring it was five, and the scan beside it cost 9 ms per 20 000 frames, r2z)
  static constexpr int TS = CH == 32 ? CH : CH + 4;
  static constexpr int ring_bytes = NG * TMA_G * SLOT;
  static constexpr int tile_off = ring_bytes;
  static constexpr int tile_bytes = TMA_NB * TS * 4;
  static constexpr int rowp_off = tile_off + tile_bytes;
  static constexpr int y_off = rowp_off + TMA_NB * 8;
  static constexpr int org_off = y_off + TMA_S * TMA_TH * 8;
  static constexpr int bar_off = org_off + TMA_S * 8;
  static constexpr int flag_off = bar_off + 2 * NG * 8;
  static constexpr int total = flag_off + TMA_S + 16;
};

// VAL1: every projection value is exactly 1.0 (one camera: psp_process.cpp:318-322), so a node-frame
// value is the integer pixel value itself and the sums are taken in integer arithmetic (exact; the
// reference's double sums of these floats are exact too, so the bits agree).
// Boxes: the G = 4 frames of a group are fetched with ONE box [G frames x BH rows] (tmapG) when the boxes of the four
// frames fit a common origin (camera shake of a few pixels between neighbouring frames), with one box per frame
// (tmap1) otherwise.  The TMA unit's cost is per box far more than per byte (measured: 80 cycles per 1 KB box per SM,
// 156 per 4.6 KB box), and with one box per frame the kernel waited on it.
// NG groups in the ring, the producer LA groups ahead of the consumers, MINB resident blocks per SM asked of ptxas.
// IT16 (with VAL1): the node-major rows are stored as 16-bit integers.  With unit projection values a node-frame value
// IS the rounded warped pixel (an integer < 2^16, exact in either type), so the row store, the all-to-all over NVLink
// and the read of phase 2 move half the bytes; the ABI's readers widen to float (upsp_gpu.cu).  The staging tile keeps its
// byte geometry: a chunk is 2 CH frames of 2 bytes instead of CH frames of 4.
template <int SRC, int CH, bool VAL1, int NG = TMA_NG, int LA = TMA_LA, int MINB = 6, bool IT16 = false>
__global__ void __launch_bounds__(TMA_NB, MINB)
k_project_tma(const __grid_constant__ CUtensorMap tmapG, const __grid_constant__ CUtensorMap tmap1, const FusedArgs a,
              const TmaExtra ex) {
  constexpr int TMA_NG = NG, TMA_LA = LA;      // shadow the defaults below
  static_assert(LA >= 1 && LA < NG, "look-ahead must leave a free group");
  static_assert(!IT16 || VAL1, "16-bit rows need integer node values");
  constexpr int CHF = IT16 ? 2 * CH : CH;      // frames per chunk
  constexpr int ESZ = IT16 ? 2 : 4;            // bytes per stored value
  static_assert(CHF <= TMA_S && TMA_S % CHF == 0, "a table stage is a whole number of chunks");
  using L = TmaSmem<SRC, CH, TMA_NG>;
  constexpr int SLOT = L::SLOT;
  constexpr int TS = L::TS;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  unsigned char* tile = smem + L::tile_off;                     // [node][TS * 4 bytes]: CHF values + 16 bytes of padding
  unsigned char** rowp = reinterpret_cast<unsigned char**>(smem + L::rowp_off);
  const double2* __restrict__ coef = c_tma_coef[ex.coef_set];   // [batch] (constant bank)
  int2* s_y = reinterpret_cast<int2*>(smem + L::y_off);        // [S][TH]: (X0,Y0)[ymin+r] minus the box origin
  int2* s_org = reinterpret_cast<int2*>(smem + L::org_off);    // [S]: box origin (px, row) of the frame
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::bar_off);
  uint64_t* empty = full + TMA_NG;
  unsigned char* s_flag = smem + L::flag_off;                  // [S]: 1 = box does not fit (slow), 2 = hot fix in the box

  const FusedCam& cam = a.cam[0];
  const TmaBlock bd = ex.blk[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool live = tid < bd.count;
  const int n = live ? __ldg(a.perm + bd.node0 + tid) : -1;
  if (IT16) {       // 16-bit rows: same addressing as fused_row_ptr in 2-byte elements (no staging in this mode)
    unsigned char* rp = nullptr;
    if (live) {
      int r = 0;
      while (r + 1 < a.n_ranks && n >= a.node_start[r + 1]) ++r;
      if (a.blk_len > 0)      // batch-blocked destination: [block][N_r][blk_len]
        rp = reinterpret_cast<unsigned char*>(a.dst[r]) +
             (((size_t)a.blk_index * (size_t)(a.node_start[r + 1] - a.node_start[r]) + (size_t)(n - a.node_start[r])) * a.blk_len + a.blk_j0) * 2;
      else
        rp = reinterpret_cast<unsigned char*>(a.dst[r]) + ((size_t)(n - a.node_start[r]) * a.f_total + a.col0) * 2;
    }
    rowp[tid] = rp;
  } else {
    rowp[tid] = live ? reinterpret_cast<unsigned char*>(fused_row_ptr(a, n)) : nullptr;
  }
  if (tid == 0) {
    for (int i = 0; i < TMA_NG; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, TMA_NB / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int W = cam.W, H = cam.H;
  const int code = live ? __ldg(cam.code + n) : bd.ymin * W + bd.xmin;     // dead lanes compute a valid pixel
  const float val = live ? __ldg(cam.val + n) : 0.0f;
  const int px = code % W, py = code / W;
  const double dpx = (double)px;
  const int yrel = py - bd.ymin;
  const bool vec_ok = a.blk_len > 0 ? ((a.blk_len | a.blk_j0) & 7) == 0
                                    : ((a.f_total | a.col0) & (IT16 ? 7 : 3)) == 0;      // 16-byte aligned row segments
  unsigned char* trow = tile + tid * (TS * 4);
  constexpr bool SWZ = CH == 32;
  // byte offset `off` inside this thread's tile row -> its place (16-byte chunks swizzled with the row index)
  auto tpos = [&](int row, int off) -> int { return SWZ ? ((((off >> 4) ^ (row & 7)) << 4) | (off & 15)) : off; };
  double s = 0.0, q = 0.0;
  unsigned gidx = 0;        // groups issued / consumed so far by this block (ring position)

  // blockIdx.y = which slice of the batch's frames (ex.split_frames each, a multiple of the table stage): twice the
  // blocks at half the length, so that the last wave of the grid is short (3953 blocks on 148 x 6 slots were 4.45 waves)
  const int f_begin = ex.split_frames > 0 ? (int)blockIdx.y * ex.split_frames : 0;
  const int f_end = ex.split_frames > 0 ? min(a.nframes, f_begin + ex.split_frames) : a.nframes;
  for (int s0 = f_begin; s0 < f_end; s0 += TMA_S) {
    const int ns = min(TMA_S, f_end - s0);
    __syncthreads();       // previous stage fully consumed (tables, ring); first pass: barriers initialised
    if (tid < TMA_S) {      // warps 0 .. S/32 - 1, whole warps: the four frames of a group are neighbouring lanes
      const bool valid = tid < ns;
      const int f = s0 + min(tid, ns - 1);
      const double2 cf = coef[f];   // identity (1024, 0) for the unregistered frame
      int2 ye[TMA_TH];
      if (f == a.skip_frame) {      // global frame 0 is never registered: identity map, exact taps
#pragma unroll
        for (int r = 0; r < TMA_TH; ++r) ye[r] = make_int2(16, (min(bd.ymin + r, H - 1) << 10) + 16);
      } else {
        const int2* ty = reinterpret_cast<const int2*>(cam.tab) + ((size_t)f * (unsigned)(W + H) + (unsigned)W);
#pragma unroll
        for (int r = 0; r < TMA_TH; ++r) ye[r] = __ldg(ty + min(bd.ymin + r, H - 1));
      }
      // X(x,y) = cvRound(M0 x 1024) + X0[y] is monotone in x and in y (same for Y): extremal at the corners
      const int axl = __double2int_rn(__dmul_rn(cf.x, (double)bd.xmin)), axh = __double2int_rn(__dmul_rn(cf.x, (double)bd.xmax));
      const int bxl = __double2int_rn(__dmul_rn(cf.y, (double)bd.xmin)), bxh = __double2int_rn(__dmul_rn(cf.y, (double)bd.xmax));
      const int2 ylo = ye[0];
      int2 yhi = ye[0];
#pragma unroll
      for (int r = 1; r < TMA_TH; ++r)
        if (r == bd.ymax - bd.ymin) yhi = ye[r];
      int sxmin = 0x7fffffff, sxmax = -0x7fffffff, symin = 0x7fffffff, symax = -0x7fffffff;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int sx = (((k & 1) ? axh : axl) + ((k & 2) ? yhi.x : ylo.x)) >> 10;
        const int sy = (((k & 1) ? bxh : bxl) + ((k & 2) ? yhi.y : ylo.y)) >> 10;
        sxmin = min(sxmin, sx);
        sxmax = max(sxmax, sx);
        symin = min(symin, sy);
        symax = max(symax, sy);
      }
      // the TMA unit wants the box to start on a 16-byte boundary of global memory (measured: any other inner
      // coordinate raises "illegal instruction"): 8 px of a u16 row, 32 px (48 bytes) of a packed row
      constexpr int XMASK = SRC ? ~31 : ~7;
      constexpr int BWPX = SRC ? (TMA_BWB12 * 2) / 3 : TMA_BW16;
      // sane coordinates only (the shifts below must not overflow); anything else takes the slow path
      const bool sane = sxmin > -(1 << 20) && sxmax < (1 << 20) && symin > -(1 << 20) && symax < (1 << 20);
      int bx0 = sxmin & XMASK, by0 = symin;
      const bool fit1 = sane && (sxmax + 1 - bx0 < BWPX) && (symax + 1 - by0 < TMA_BH);
      // common box of the group's four frames?
      int gxmin = sxmin, gxmax = sxmax, gymin = symin, gymax = symax;
      bool gall = valid && sane;
#pragma unroll
      for (int d = 1; d < TMA_G; d <<= 1) {
        gxmin = min(gxmin, __shfl_xor_sync(0xffffffffu, gxmin, d));
        gxmax = max(gxmax, __shfl_xor_sync(0xffffffffu, gxmax, d));
        gymin = min(gymin, __shfl_xor_sync(0xffffffffu, gymin, d));
        gymax = max(gymax, __shfl_xor_sync(0xffffffffu, gymax, d));
        gall = __shfl_xor_sync(0xffffffffu, (int)gall, d) && gall;
      }
      const int gbx0 = gxmin & XMASK;
      const bool gfit = gall && (gxmax + 1 - gbx0 < BWPX) && (gymax + 1 - gymin < TMA_BH);
      if (gfit) {
        bx0 = gbx0;
        by0 = gymin;
      }
      const bool fit = gfit || fit1;
      unsigned char flag = (fit ? 0 : 1) | (gfit ? 4 : 0);
      if (SRC && fit && valid && ex.hot != nullptr) {
        const HotFix* h = ex.hot + f;
        const int nh = h->n;
        for (int i = 0; i < nh; ++i) {
          const int hx = h->pos[i] % W - bx0, hy = h->pos[i] / W - by0;
          if ((unsigned)hx < (unsigned)BWPX && (unsigned)hy < (unsigned)TMA_BH) flag |= 2;
        }
      }
      if (valid) {
#pragma unroll
        for (int r = 0; r < TMA_TH; ++r) s_y[tid * TMA_TH + r] = make_int2(ye[r].x - (fit ? bx0 << 10 : 0), ye[r].y - (fit ? by0 << 10 : 0));
        s_org[tid] = make_int2(fit ? bx0 : 0, fit ? by0 : 0);
        s_flag[tid] = flag;
      } else {
        s_flag[tid] = 0;
      }
    }
    __syncthreads();
    const int ngr = (ns + TMA_G - 1) / TMA_G;
    // producer side (thread 0): fill ring position `gi` with group g of this stage
    auto issue = [&](int g, unsigned gi) {
      const unsigned slot = gi % TMA_NG, use = gi / TMA_NG;
      if (use > 0) mbar_wait(empty + slot, (use - 1) & 1);
      const int nf = min(TMA_G, ns - g * TMA_G);
      mbar_expect_tx(full + slot, (uint32_t)nf * SLOT);
      if (s_flag[g * TMA_G] & 4) {       // one box for the group's four frames
        const int2 o = s_org[g * TMA_G];
        tma_load_3d(ring + (slot * TMA_G) * SLOT, &tmapG, full + slot, SRC ? (o.x >> 5) * 12 : o.x, o.y, ex.frame0 + s0 + g * TMA_G);
        return;
      }
      for (int j = 0; j < nf; ++j) {
        const int i = g * TMA_G + j;
        const int2 o = s_org[i];
        tma_load_3d(ring + (slot * TMA_G + j) * SLOT, &tmap1, full + slot, SRC ? (o.x >> 5) * 12 : o.x, o.y, ex.frame0 + s0 + i);
      }
    };
    if (tid == 0)
      for (int g = 0; g < min(TMA_LA, ngr); ++g) issue(g, gidx + g);

    for (int c0 = 0; c0 < ns; c0 += CHF) {
      const int nb = min(CHF, ns - c0);
      const int b0 = s0 + c0;
      for (int u = 0; u < nb; u += TMA_G) {
        const int g = (c0 + u) / TMA_G;
        const unsigned gi = gidx + g;
        const unsigned slot = gi % TMA_NG;
        if (tid == 0 && g + TMA_LA < ngr) issue(g + TMA_LA, gi + TMA_LA);
        const int nf = min(TMA_G, nb - u);
        const unsigned flags = *reinterpret_cast<const unsigned*>(s_flag + c0 + u);    // 4 frames' flags (c0+u is a multiple of 4)
        mbar_wait(full + slot, (gi / TMA_NG) & 1);
        if (SRC && (flags & 0x02020202u)) {
          // hot-pixel fixes inside the staged boxes: patch the packed bytes (block-uniform branch)
          __syncthreads();
          if (tid < nf && ((flags >> (8 * tid)) & 2u)) {
            const int i = c0 + u + tid;
            const HotFix* h = ex.hot + (s0 + i);
            const int2 o = s_org[i];
            unsigned char* sl = ring + (slot * TMA_G + tid) * SLOT;
            for (int k = 0; k < h->n; ++k) {
              const int hx = h->pos[k] % W - o.x, hy = h->pos[k] / W - o.y;
              if ((unsigned)hx < (unsigned)((TMA_BWB12 * 2) / 3) && (unsigned)hy < (unsigned)TMA_BH) {
                unsigned char* p = sl + hy * TMA_BWB12 + (hx >> 1) * 3;
                const unsigned v = (unsigned)h->val[k];
                if (hx & 1) {
                  p[1] = (unsigned char)((p[1] & 0xF0u) | (v >> 8));
                  p[2] = (unsigned char)v;
                } else {
                  p[0] = (unsigned char)(v >> 4);
                  p[1] = (unsigned char)((p[1] & 0x0Fu) | ((v & 0xFu) << 4));
                }
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          }
          __syncthreads();
        }
        float sol[TMA_G];
        unsigned ri[TMA_G];
        if (nf == TMA_G && !(flags & 0x01010101u)) {
#pragma unroll
          for (int j = 0; j < TMA_G; ++j) {
            const int i = c0 + u + j;
            const double2 cf = coef[s0 + i];
            const int2 ya = s_y[i * TMA_TH + yrel];
            const int X = ya.x + __double2int_rn(__dmul_rn(cf.x, dpx));
            const int Y = ya.y + __double2int_rn(__dmul_rn(cf.y, dpx));
            const float t = tma_px<SRC>(ring + (slot * TMA_G + j) * SLOT, X, Y);
            ri[j] = __float_as_uint(t) - 0x4B400000u;
            const float v = __fadd_rn(t, -12582912.0f);
            sol[j] = VAL1 ? v : __fmaf_rn(val, v, 0.0f);
          }
        } else {
          // tail group of the stage, or a frame whose box does not fit (every tap through global memory)
#pragma unroll
          for (int j = 0; j < TMA_G; ++j) {
            sol[j] = 0.0f;
            ri[j] = 0u;
            if (j >= nf) continue;
            const int i = c0 + u + j;
            const double2 cf = coef[s0 + i];
            const int2 ya = s_y[i * TMA_TH + yrel];
            const int X = ya.x + __double2int_rn(__dmul_rn(cf.x, dpx));
            const int Y = ya.y + __double2int_rn(__dmul_rn(cf.y, dpx));
            float v;
            if ((flags >> (8 * j)) & 1u) {
              if (SRC == 0) v = warp_px_slow(cam.frames + (size_t)(s0 + i) * cam.npix, W, H, X, Y, 1);
              else v = warp_px_slow12(ex.packed + (size_t)(s0 + i) * ex.frame_bytes, ex.hot ? ex.hot + (s0 + i) : nullptr, W, H, X, Y);
            } else {
              v = __fadd_rn(tma_px<SRC>(ring + (slot * TMA_G + j) * SLOT, X, Y), -12582912.0f);
            }
            ri[j] = (unsigned)(int)v;
            sol[j] = VAL1 ? v : __fmaf_rn(val, v, 0.0f);
          }
        }
        // this warp is done with the ring slot
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
        if (nf == TMA_G) {
          if (IT16) *reinterpret_cast<uint2*>(trow + tpos(tid, u * 2)) = make_uint2(ri[0] | (ri[1] << 16), ri[2] | (ri[3] << 16));
          else *reinterpret_cast<float4*>(trow + tpos(tid, u * 4)) = make_float4(sol[0], sol[1], sol[2], sol[3]);
          if (VAL1) {
            unsigned si = 0, qi = 0;      // 4 * 4095 and 4 * 4095^2 fit easily
#pragma unroll
            for (int j = 0; j < TMA_G; ++j) {
              si += ri[j];
              qi += ri[j] * ri[j];
            }
            s += (double)si;
            q += (double)qi;
          } else {
#pragma unroll
            for (int j = 0; j < TMA_G; ++j) {
              q += (double)__fmul_rn(sol[j], sol[j]);
              s += (double)sol[j];
            }
          }
        } else {
          for (int j = 0; j < nf; ++j) {
            if (IT16) *reinterpret_cast<unsigned short*>(trow + tpos(tid, (u + j) * 2)) = (unsigned short)ri[j];
            else *reinterpret_cast<float*>(trow + tpos(tid, (u + j) * 4)) = sol[j];
            q += (double)__fmul_rn(sol[j], sol[j]);
            s += (double)sol[j];
          }
        }
      }
      __syncwarp();
      if (vec_ok && nb == CHF) {
        constexpr int LPN = CH / 4;      // lanes per node row segment (16 bytes each)
        const int fq = (lane % LPN) * 16;      // byte offset inside the segment
        // CH = 16: the two nodes of a quarter warp are 4 tile rows apart (4 * TS = 80 words = 16 banks: their two
        // 16-word segments tile the 32 banks; neighbouring rows, 20 banks apart, overlapped in 4 of them)
        const int nsel = CH == 16 ? (lane >> 3) + 4 * ((lane >> 2) & 1) : lane / LPN;
#pragma unroll
        for (int it = 0; it < LPN; ++it) {
          const int nl = w * 32 + it * (32 / LPN) + nsel;
          unsigned char* rp = rowp[nl];
          if (rp != nullptr) {
            const float4 o = *reinterpret_cast<const float4*>(tile + nl * (TS * 4) + tpos(nl, fq));
            *reinterpret_cast<float4*>(rp + (size_t)b0 * ESZ + fq) = o;
          }
        }
      } else {
        for (int j = 0; j < 32; ++j) {
          unsigned char* rp = rowp[w * 32 + j];
          if (rp == nullptr) continue;
          const unsigned char* tr = tile + (w * 32 + j) * (TS * 4);
          for (int f = lane; f < nb; f += 32) {
            if (IT16) reinterpret_cast<unsigned short*>(rp)[b0 + f] = *reinterpret_cast<const unsigned short*>(tr + tpos(w * 32 + j, f * 2));
            else reinterpret_cast<float*>(rp)[b0 + f] = *reinterpret_cast<const float*>(tr + tpos(w * 32 + j, f * 4));
          }
        }
      }
      __syncwarp();
    }
    gidx += ngr;
  }
  if (live) {
    if (ex.split_frames > 0) {      // several blocks per node: VAL1 only (integer sums, exact in any order)
      atomicAdd(a.sum + n, s);
      atomicAdd(a.sumsq + n, q);
    } else {
      a.sum[n] += s;
      a.sumsq[n] += q;
    }
  }
}

// ---- SRC 1 front end: the scan half of fix_hot_pixels on the PACKED frames (read only) and the
// <= 5 fix-ups of a frame as a (position, new value) list.  Reference: cpp/utils/cv_extras.cpp:230-272
// (>= 4064 scan, more than 5 hot pixels: frame untouched; raster order, a later fix sees an earlier one;
// median of the 4-neighbourhood = vals[n/2] after sort; replace if the drop exceeds 512).
// One thread-iteration = 48 packed bytes = 32 pixels; persistent grid; the block that completes a frame
// (item counter) builds the list.
static __device__ __noinline__ void hot_list_build(const uint8_t* __restrict__ fr, int rows, int cols, int n, const int* __restrict__ pos,
                                            HotFix* __restrict__ out) {
  int loc[UPSP_HOT_STORE];
  for (int i = 0; i < n; ++i) loc[i] = pos[i];
  for (int i = 1; i < n; ++i) {
    int v = loc[i], j = i - 1;
    while (j >= 0 && loc[j] > v) {
      loc[j + 1] = loc[j];
      --j;
    }
    loc[j + 1] = v;
  }
  HotFix h;
  h.n = 0;
  h.pad = 0;
  for (int i = 0; i < UPSP_HOT_MAX; ++i) h.pos[i] = h.val[i] = 0;
  auto get = [&](int idx) -> int {
    for (int i = 0; i < h.n; ++i)
      if (h.pos[i] == idx) return h.val[i];
    return (int)px_packed12(fr, (unsigned)idx);
  };
  for (int k = 0; k < n; ++k) {
    const int row = loc[k] / cols, col = loc[k] % cols;
    int vals[4], nv = 0;
    if (row > 0) vals[nv++] = get((row - 1) * cols + col);
    if (col > 0) vals[nv++] = get(row * cols + col - 1);
    if (row < rows - 1) vals[nv++] = get((row + 1) * cols + col);
    if (col < cols - 1) vals[nv++] = get(row * cols + col + 1);
    for (int i = 1; i < nv; ++i) {
      int v = vals[i], j = i - 1;
      while (j >= 0 && vals[j] > v) {
        vals[j + 1] = vals[j];
        --j;
      }
      vals[j + 1] = v;
    }
    const int old_val = get(loc[k]);
    const int new_val = vals[nv / 2];
    if (old_val - new_val > UPSP_HOT_MIN_CHANGE) {
      h.pos[h.n] = loc[k];
      h.val[h.n] = new_val;
      h.n++;
    }
  }
  *out = h;
}

// rare path of the scan: one item (48 packed bytes = 32 pixels) holds a hot pixel; re-read it and note which
static __device__ __noinline__ void note_hot_item12(const uint8_t* __restrict__ frame, unsigned i, int thresh, int* cnt, int* pos) {
  for (unsigned k = 0; k < 32; ++k) note_hot(px_packed12(frame, i * 32 + k), (size_t)i * 32 + k, thresh, cnt, pos);
}

// Small on purpose (<= 32 registers, any block size): it runs on the front-end stream UNDER the projection of
// the previous batch, in the registers / issue slots that kernel leaves free.
__global__ void __launch_bounds__(256, 8)
k_hot_scan12(const uint8_t* __restrict__ in, size_t in_stride, size_t npix, int nframes, int thresh,
             int* __restrict__ hot_cnt, int* __restrict__ hot_pos, int* __restrict__ done, int rows, int cols,
             HotFix* __restrict__ fixes) {
  const unsigned ipf = (unsigned)(npix / 32);
  const unsigned total = ipf * (unsigned)nframes;
  const unsigned chunk = (total + gridDim.x - 1) / gridDim.x;
  unsigned it0 = blockIdx.x * chunk;
  const unsigned it1 = min(total, it0 + chunk);
  // a 12-bit pixel is >= thresh iff ... tested on the unpacked pair words like k_unpack12_scan_p
  const uint32_t t2 = (uint32_t)min(thresh, 0x8000) * 0x00010001u;
  while (it0 < it1) {
    const unsigned f = it0 / ipf;
    const unsigned fbeg = f * ipf;
    const unsigned seg_end = min(it1, fbeg + ipf);
    const uint8_t* frame = in + (size_t)f * in_stride;
    const uint4* src = reinterpret_cast<const uint4*>(frame);
    const unsigned iend = seg_end - fbeg;
    for (unsigned i = it0 - fbeg + threadIdx.x; i < iend; i += blockDim.x) {
      const uint4 b0 = ld_stream_u4(src + 3 * i), b1 = ld_stream_u4(src + 3 * i + 1), b2 = ld_stream_u4(src + 3 * i + 2);
      uint32_t hot = 0;
      uint4 o;
#define UPSP_SCAN_X8(W0, W1, W2)                                                                              \
  unpack12_x8(W0, W1, W2, o);                                                                                 \
  hot |= (((o.x | 0x80008000u) - t2) | ((o.y | 0x80008000u) - t2) | ((o.z | 0x80008000u) - t2) | ((o.w | 0x80008000u) - t2))
      UPSP_SCAN_X8(b0.x, b0.y, b0.z);
      UPSP_SCAN_X8(b0.w, b1.x, b1.y);
      UPSP_SCAN_X8(b1.z, b1.w, b2.x);
      UPSP_SCAN_X8(b2.y, b2.z, b2.w);
#undef UPSP_SCAN_X8
      if (hot & 0x80008000u) note_hot_item12(frame, i, thresh, hot_cnt + f, hot_pos + f * UPSP_HOT_STORE);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int items = (int)(seg_end - it0);
      if (atomicAdd(done + f, items) + items == (int)ipf) {
        __threadfence();
        const int nh = *((volatile int*)(hot_cnt + f));
        if (nh > 0 && nh <= UPSP_HOT_MAX) hot_list_build(frame, rows, cols, nh, hot_pos + f * UPSP_HOT_STORE, fixes + f);
        else fixes[f].n = 0;
      }
    }
    it0 = seg_end;
  }
}

}  // namespace upsp
