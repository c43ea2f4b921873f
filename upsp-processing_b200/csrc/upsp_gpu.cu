// upsp_gpu.cu -- C ABI (include/upsp_gpu.h) over the sm_100a kernels.
// Context lifecycle, device memory layout, batching, stream/event plumbing, multi-GPU
// wiring.  No CPU fallback anywhere: every compute entry point launches CUDA kernels.
#include "../../include/upsp_gpu.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "host_qr.hpp"
#include "kernels_frame.cuh"
#include "kernels_ecc.cuh"
#include "kernels_filter.cuh"
#include "kernels_phase2.cuh"
#include "kernels_project.cuh"
#include "kernels_setup.cuh"
#include "kernels_transpose.cuh"
#include "proj_tma.h"

using namespace upsp;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess)                                                            \
      return fail(e_ == cudaErrorMemoryAllocation ? UPSP_ERR_NOMEM : UPSP_ERR_CUDA,   \
                  "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

#define REQUIRE(cond, code, ...) \
  do {                           \
    if (!(cond)) return fail(code, __VA_ARGS__); \
  } while (0)

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// apportion() cpp/exec/psp_process.cpp:611-624
static void apportion(int value, int bins, std::vector<int>& start, std::vector<int>& extent) {
  start.assign(bins, 0);
  extent.assign(bins, 0);
  long block = value / bins, rem = value - block * bins, next = 0;
  for (int b = 0; b < bins; ++b) {
    start[b] = (int)next;
    extent[b] = (int)(block + (b < rem));
    next += extent[b];
  }
}

// ------------------------------------------------------------------------------------------
// Shareable device memory (CUDA VMM).  Legacy cudaIpc mappings of cudaMalloc memory are slow
// for kernel-issued peer stores on these boxes (measured: 5-40x slower than the same kernel
// over cudaDeviceEnablePeerAccess mappings), so the cross-process shared block is a
// cuMemCreate allocation exported as a POSIX file descriptor; peers duplicate the descriptor
// with pidfd_getfd and map it with cuMemMap (2 MB pages, NVLink peer access).
// The driver entry points are resolved at run time: no link-time dependency on libcuda.
// ------------------------------------------------------------------------------------------
struct DriverApi {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
  bool ok = false;
};

static DriverApi& drv() {
  static DriverApi d;
  static bool tried = false;
  if (!tried) {
    tried = true;
    bool ok = true;
    auto get = [&](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || *fn == nullptr) ok = false;
    };
    get("cuMemCreate", (void**)&d.MemCreate);
    get("cuMemRelease", (void**)&d.MemRelease);
    get("cuMemAddressReserve", (void**)&d.MemAddressReserve);
    get("cuMemAddressFree", (void**)&d.MemAddressFree);
    get("cuMemMap", (void**)&d.MemMap);
    get("cuMemUnmap", (void**)&d.MemUnmap);
    get("cuMemSetAccess", (void**)&d.MemSetAccess);
    get("cuMemExportToShareableHandle", (void**)&d.MemExportToShareableHandle);
    get("cuMemImportFromShareableHandle", (void**)&d.MemImportFromShareableHandle);
    get("cuMemGetAllocationGranularity", (void**)&d.MemGetAllocationGranularity);
    d.ok = ok;
    {   // TMA descriptors (projection kernels); absence only disables those kernels
      cudaDriverEntryPointQueryResult q;
      void* fn = nullptr;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn != nullptr)
        d.TensorMapEncodeTiled = reinterpret_cast<decltype(d.TensorMapEncodeTiled)>(fn);
    }
  }
  return d;
}

// ------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen of the libnccl.so.2 the process already has, e.g. torch's): the library has no
// link-time dependency on it, and only the UPSP_XCHG_NCCL exchange needs it.  Minimal declarations of nccl.h.
// ------------------------------------------------------------------------------------------
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  typedef void* Comm;
  void* so = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  static constexpr int kFloat = 7, kDouble = 8, kSum = 0;      // ncclFloat32, ncclFloat64, ncclSum
};

static NcclApi& nccl() {
  static NcclApi a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  const char* names[] = {getenv("UPSP_NCCL_LIB"), "libnccl.so.2", "libnccl.so",
                         "/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/lib/libnccl.so.2"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    a.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.so) break;
  }
  if (!a.so) return a;
#define UPSP_NCCL_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.so, name))
  UPSP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  UPSP_NCCL_SYM(CommInitRank, "ncclCommInitRank");
  UPSP_NCCL_SYM(CommDestroy, "ncclCommDestroy");
  UPSP_NCCL_SYM(GroupStart, "ncclGroupStart");
  UPSP_NCCL_SYM(GroupEnd, "ncclGroupEnd");
  UPSP_NCCL_SYM(Send, "ncclSend");
  UPSP_NCCL_SYM(Recv, "ncclRecv");
  UPSP_NCCL_SYM(AllReduce, "ncclAllReduce");
  UPSP_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef UPSP_NCCL_SYM
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Send && a.Recv && a.AllReduce;
  return a;
}
#define NCCLCHK(expr)                                                                                   \
  do {                                                                                                  \
    int r_ = (expr);                                                                                    \
    if (r_ != 0) return fail(UPSP_ERR_COMM, "%s -> %s", #expr, nccl().GetErrorString ? nccl().GetErrorString(r_) : "?"); \
  } while (0)

struct VmmBlock {
  CUmemGenericAllocationHandle handle = 0;
  CUdeviceptr ptr = 0;
  size_t size = 0;
  int fd = -1;
  bool mapped = false;
};

static int vmm_map(VmmBlock& b, int device) {
  DriverApi& d = drv();
  CUresult r = d.MemAddressReserve(&b.ptr, b.size, 0, 0, 0);
  if (r != CUDA_SUCCESS) return fail(UPSP_ERR_NOMEM, "cuMemAddressReserve(%zu) -> %d", b.size, (int)r);
  r = d.MemMap(b.ptr, b.size, 0, b.handle, 0);
  if (r != CUDA_SUCCESS) return fail(UPSP_ERR_NOMEM, "cuMemMap -> %d", (int)r);
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.MemSetAccess(b.ptr, b.size, &acc, 1);
  if (r != CUDA_SUCCESS) return fail(UPSP_ERR_COMM, "cuMemSetAccess(device %d) -> %d", device, (int)r);
  b.mapped = true;
  return UPSP_OK;
}

static int vmm_alloc(VmmBlock& b, size_t bytes, int device) {
  DriverApi& d = drv();
  if (!d.ok) return fail(UPSP_ERR_COMM, "CUDA VMM driver entry points unavailable");
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  CUresult r = d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS || gran == 0) return fail(UPSP_ERR_CUDA, "cuMemGetAllocationGranularity -> %d", (int)r);
  b.size = (bytes + gran - 1) / gran * gran;
  r = d.MemCreate(&b.handle, b.size, &prop, 0);
  if (r != CUDA_SUCCESS) return fail(UPSP_ERR_NOMEM, "cuMemCreate(%zu bytes) -> %d", b.size, (int)r);
  int rc = vmm_map(b, device);
  if (rc) return rc;
  r = d.MemExportToShareableHandle(&b.fd, b.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) return fail(UPSP_ERR_COMM, "cuMemExportToShareableHandle -> %d", (int)r);
  return UPSP_OK;
}

static void vmm_free(VmmBlock& b) {
  DriverApi& d = drv();
  if (b.mapped) {
    d.MemUnmap(b.ptr, b.size);
    d.MemAddressFree(b.ptr, b.size);
  }
  if (b.handle) d.MemRelease(b.handle);
  if (b.fd >= 0) close(b.fd);
  b = VmmBlock();
}

struct IpcWire {          // the 64-byte handle exchanged by the host
  uint32_t magic;         // 'UPSV'
  int32_t pid;
  int32_t fd;
  int32_t device;
  uint64_t size;
  char pad[40];
};
static_assert(sizeof(IpcWire) == UPSP_IPC_HANDLE_BYTES, "handle size");

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct Camera {
  int W = 0, H = 0;
  size_t npix = 0;
  // caller's CSR (kept on the host until finalize folds remap + patch codes into it)
  std::vector<int> rowptr, col;
  std::vector<float> val;
  bool has_proj = false;
  // input store (device-resident slots) and per-batch working buffers
  int format = -1;
  size_t frame_bytes = 0;
  uint8_t* d_in = nullptr;
  uint16_t* d_work = nullptr;
  uint16_t* d_warp = nullptr;
  int* d_hot_cnt = nullptr;
  int* d_hot_pos = nullptr;
  // registration
  float* d_m6 = nullptr;      // [F_local][6]
  float* d_rho = nullptr;     // [F_local]
  int* d_iters = nullptr;     // [F_local]
  int* d_tab = nullptr;       // [F_local][2W+2H] warp tables of every local frame
  double2* d_coef = nullptr;  // [F_local] (M0, M3) * 1024 of every local frame, identity for global frame 0 (TMA projection)
  uint8_t* d_skip = nullptr;  // [batch]
  uint16_t* d_ref16 = nullptr;
  // ECC (registration = pixel)
  float *d_eccT = nullptr, *d_eccI = nullptr, *d_eccTmp = nullptr;
  float2* d_eccG = nullptr;
  double* d_eccPart = nullptr;
  EccState* d_eccState = nullptr;   // [F_local]
  int* d_eccTab = nullptr;
  int* d_nactive = nullptr;
  bool has_ref = false, has_m6 = false;
  // patches
  bool has_patches = false;
  PatchGeom geom{};
  std::vector<void*> patch_allocs;
  std::vector<std::vector<int>> levels;  // cluster ids per dependency level
  int* d_cl_list = nullptr;              // concatenated level lists
  std::vector<int> level_off;
  int total_bounds = 0, total_internal = 0, max_bounds = 0;
  float* d_scratch = nullptr;
  float* d_pv = nullptr;
  // two-stream pipeline: second set of per-batch buffers (set 0 = the members above)
  uint16_t* d_work2 = nullptr;
  int* d_hot_cnt2 = nullptr;
  int* d_hot_pos2 = nullptr;
  float* d_pv2 = nullptr;
  std::unordered_map<int, int> pix2slot;  // interior pixel -> slot of the LAST active cluster
  // spatial filter (unfused mode only)
  float *d_img32 = nullptr, *d_img32b = nullptr;
  uint16_t* d_filt16 = nullptr;
  int* d_slot_pix = nullptr;   // [interior slots] pixel index, or -1 if a later cluster overwrites it
  // projection tables
  int* d_code = nullptr;
  float* d_val = nullptr;
  int* d_rowptr = nullptr;
  std::vector<int> h_code;     // ELL-1 table as uploaded (block partition of the TMA projection)
  std::vector<float> h_val;
  // TMA projection: descriptors of the two decoded work buffers / of the packed input store; fix lists per buffer set
  CUtensorMap tmap16[2], tmap16g[2];     // one-frame boxes / boxes of a whole frame group (kernels_project_tma.cuh)
  CUtensorMap tmap12, tmap12g;
  void* d_fix[2] = {nullptr, nullptr};
};

struct upsp_gpu_ctx {
  upsp_gpu_config cfg{};
  int F = 0, N = 0, R = 1, rank = 0;
  std::vector<int> f_start, f_count, n_start, n_count;
  int F_local = 0, N_local = 0, f0 = 0, n0 = 0;
  int capacity = 0, batch = 32;
  int registration = UPSP_REG_NONE, interp = UPSP_INTERP_LINEAR, patcher = UPSP_PATCH_NONE;
  int hot_fix = 1;
  FilterSpec filter{};   // kind 0 = none
  std::vector<Camera> cams;
  std::vector<int> remap;  // src_index or empty
  bool finalized = false, ell1 = true, fused = false;
  uint16_t* d_lut = nullptr;
  int lut_max = 0;        // largest entry of the 10->12-bit table (0 = no table)
  int* d_perm = nullptr;  // fused mode: node processing order (Morton order of the nodes' pixels)
  // TMA-staged projection (kernels_project_tma.cuh): 0 = off (gather kernels), 1 = boxes from the decoded u16 frames,
  // 2 = boxes from the packed 12-bit frames (no decode pass).  Decided at the first batch (needs the pixel format).
  int proj_mode = -1;
  int* d_perm_tma = nullptr;      // plain-pixel nodes in block order, then the others
  TmaBlock* d_tma_blk = nullptr;
  int n_tma_blocks = 0, n_tma_plain = 0;
  bool tma_val1 = false;

  cudaStream_t stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr;
  // front-end stream (decode + hot pixels + patch of batch i+1 while the fused projection of batch i
  // runs on `stream`); high priority so its few long-lived blocks get SM slots as soon as they free up
  cudaStream_t stream_b = nullptr;
  cudaEvent_t ev_front[2] = {nullptr, nullptr}, ev_back[2] = {nullptr, nullptr}, ev_tabs = nullptr;
  cudaEvent_t ev_dec[2] = {nullptr, nullptr};    // decode / scan (+ coefficient table) of a buffer set done: the TMA kernel may start
  bool pipelined = false;
  bool last_sampled = false;   // the previous batch ran un-overlapped for kernel timing
  long pipe_batches = 0;
  int n_sm = 148;
  // staged exchange (n_ranks > 1, pipelined): the projection writes other ranks' rows into a local
  // staging block; copy engines ship them (one strided 2-D copy per peer and batch) on stream_x
  float* d_stage[2] = {nullptr, nullptr};
  int stage_stride = 0;
  static constexpr int NX = 4;     // exchange streams: strided peer copies in flight on several copy engines
  cudaStream_t stream_x[NX] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_x[2][NX] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
  bool staged_xchg = false;
  unsigned stage_mask = 0;         // ranks whose rows travel through the staging block
  int staged_peers = 0;
  // 16-bit row mode (TMA projection with unit values, decided in ensure_proj_mode): rows of plain nodes are 16-bit
  // integers at the start of the shared block, the other (patched / unseen) nodes keep float rows in a side buffer behind them
  bool it16 = false;
  // ... and in multi-rank runs those 16-bit rows live in a batch-blocked layout [source rank * blk_kb + local batch][N_local]
  // [blk_len]: what a batch of the projection stores into a peer is then one contiguous region per peer instead of
  // 128-byte pieces of 80-320 KB rows
  int blk_len = 0, blk_kb = 0;
  int* d_other_idx = nullptr;       // [N]: -1 plain node, else row in the owner's side buffer
  int* d_other_local = nullptr;     // local indices of this rank's side-buffer nodes
  int n_other_local = 0;
  size_t side_off[UPSP_MAX_RANKS] = {0};   // byte offset of rank r's side buffer in its shared block
  double* d_cl_parts = nullptr;     // clustered phase 2: [row][CL][4] partial statistics
  size_t cl_parts_n = 0;
  float* d_p2coef = nullptr;        // streaming phase 2: coefficients (+ gain) per local row
  double* d_p2parts = nullptr;      // ... and the chunks' partial sums
  size_t p2coef_n = 0, p2parts_n = 0;
  float* d_bounce = nullptr;        // readers: rows widened to float on their way to the host
  size_t bounce_floats = 0;
  bool front_serial_now = false;     // this batch's front end runs after the previous projection (process_batch_impl)
  bool ship_sm = false;           // staged rows shipped by k_ship_rows instead of the copy engines
  int ship_bpsm = 1;
  cudaEvent_t ev_push = nullptr, ev_proc = nullptr, ev_a = nullptr, ev_b = nullptr;
  cudaEvent_t ev_pa = nullptr, ev_pb = nullptr;  // process_frames timing
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;  // user timer
  struct ProcRec { int off, count; cudaEvent_t ev; };
  std::vector<ProcRec> proc_recs;                // recent process_frames calls (input-ring reuse)
  size_t proc_next = 0;
  bool push_wait_all = false;                    // next push waits on ev_proc (set by reset_run)
  // sampled per-kernel timing
  bool timeline = false;          // bracket every kernel, keep the pipeline overlapped (upsp_gpu_timeline)
  int sample_every = 0;
  long batch_counter = 0;
  std::vector<cudaEvent_t> kev;   // event pairs
  std::vector<int> kcls;
  size_t kn = 0;
  float stage_ms[4] = {0, 0, 0, 0};
  long long launches = 0;

  // big buffers
  float* d_intensity = nullptr;  // [F_local][N]
  char* d_shared = nullptr;      // one allocation (shareable): [itrans | sum | sumsq]
  VmmBlock shared_vmm;           // n_ranks > 1: d_shared is a VMM allocation
  VmmBlock peer_vmm[UPSP_MAX_RANKS];
  float* d_itrans = nullptr;     // [N_local][F]
  double* d_sum = nullptr;       // [N]
  double* d_sumsq = nullptr;     // [N]
  float* d_ptrans = nullptr;     // [N_local][F] (may alias d_intensity)
  bool ptrans_owned = false;
  float *d_avg = nullptr, *d_rms = nullptr, *d_cov = nullptr;  // [N]
  float *d_steady = nullptr, *d_temp = nullptr;                // [N]
  double *d_rms2 = nullptr, *d_avg2 = nullptr, *d_gain2 = nullptr;     // [N_local]
  float *d_rms2f = nullptr, *d_avg2f = nullptr, *d_gain2f = nullptr;   // [N_local]
  size_t shared_bytes = 0, off_sum = 0, off_sumsq = 0;
  double *d_tsum = nullptr, *d_tsq = nullptr;   // n_ranks > 1: all-reduced sums
  double** d_peer_ptrs = nullptr;                // [2R] peers' sum / sumsq bases

  // peers: base of every rank's shared allocation as seen from this device
  char* peer_base[UPSP_MAX_RANKS] = {nullptr};
  bool peer_is_ipc[UPSP_MAX_RANKS] = {false};
  bool peers_ready = false;
  int exchange = UPSP_XCHG_PEER;
  NcclApi::Comm nccl_comm = nullptr;       // UPSP_XCHG_NCCL
  float* d_nccl_send = nullptr;            // [N][chunk] blocks per destination rank
  float* d_nccl_recv = nullptr;            // [R][N_local][chunk]
  bool phase1_done = false, transposed = false, phase2_done = false;
  int frames_processed = 0;
};

static int set_dev(const upsp_gpu_ctx* c) {
  CU(cudaSetDevice(c->cfg.device));
  return UPSP_OK;
}
#define ENTER(ctx)                                               \
  REQUIRE((ctx) != nullptr, UPSP_ERR_INVALID, "null context");   \
  {                                                              \
    int rc_ = set_dev(ctx);                                      \
    if (rc_) return rc_;                                         \
  }
#define KCHECK(ctx)            \
  do {                         \
    (ctx)->launches++;         \
    CU(cudaGetLastError());    \
  } while (0)

// sampled kernel timing: KBEGIN/KEND bracket one launch with events when `on`
static int kprof_begin(upsp_gpu_ctx* c, bool on, int cls, cudaStream_t st = nullptr) {
  if (!on) return UPSP_OK;
  if (2 * c->kn + 2 > c->kev.size()) {
    if (c->kev.size() >= 16384) return UPSP_OK;  // pool exhausted: stop sampling
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e;
      CU(cudaEventCreate(&e));
      c->kev.push_back(e);
    }
    c->kcls.push_back(cls);
  }
  c->kcls[c->kn] = cls;
  CU(cudaEventRecord(c->kev[2 * c->kn], st ? st : c->stream));
  return UPSP_OK;
}
static int kprof_end(upsp_gpu_ctx* c, bool on, cudaStream_t st = nullptr) {
  if (!on || 2 * c->kn + 2 > c->kev.size()) return UPSP_OK;
  CU(cudaEventRecord(c->kev[2 * c->kn + 1], st ? st : c->stream));
  c->kn++;
  return UPSP_OK;
}
#define KBEGIN(cls) TRY(kprof_begin(c, prof, cls))
#define KEND() TRY(kprof_end(c, prof))
#define KBEGIN_ON(cls, st) TRY(kprof_begin(c, prof, cls, st))
#define KEND_ON(st) TRY(kprof_end(c, prof, st))

template <typename T>
static int dmalloc(T** p, size_t count) {
  CU(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
  return UPSP_OK;
}
template <typename T>
static int upload(T** p, const T* h, size_t count) {
  int rc = dmalloc(p, count);
  if (rc) return rc;
  if (count) CU(cudaMemcpy(*p, h, count * sizeof(T), cudaMemcpyHostToDevice));
  return UPSP_OK;
}
#define TRY(x)            \
  do {                    \
    int rc__ = (x);       \
    if (rc__) return rc__; \
  } while (0)

// ------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------
extern "C" const char* upsp_gpu_last_error(void) { return g_err.c_str(); }

extern "C" int upsp_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int upsp_gpu_create(const upsp_gpu_config* cfg, upsp_gpu_ctx** out) {
  REQUIRE(cfg && out, UPSP_ERR_INVALID, "null argument");
  REQUIRE(cfg->n_cams >= 1 && cfg->n_cams <= UPSP_MAX_CAMS, UPSP_ERR_INVALID,
          "n_cams must be in [1,%d]", UPSP_MAX_CAMS);
  REQUIRE(cfg->n_nodes > 0 && cfg->n_frames_total > 0, UPSP_ERR_INVALID, "empty problem");
  REQUIRE(cfg->n_ranks >= 1 && cfg->n_ranks <= UPSP_MAX_RANKS && cfg->rank >= 0 &&
              cfg->rank < cfg->n_ranks,
          UPSP_ERR_INVALID, "bad rank %d / %d", cfg->rank, cfg->n_ranks);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  REQUIRE(e == cudaSuccess && ndev > 0, UPSP_ERR_CUDA,
          "no usable CUDA device (%s); libupsp_gpu has no CPU fallback",
          e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
  REQUIRE(cfg->device >= 0 && cfg->device < ndev, UPSP_ERR_INVALID, "device %d of %d",
          cfg->device, ndev);
  auto* c = new upsp_gpu_ctx();
  c->cfg = *cfg;
  c->F = cfg->n_frames_total;
  c->N = cfg->n_nodes;
  c->R = cfg->n_ranks;
  c->rank = cfg->rank;
  apportion(c->F, c->R, c->f_start, c->f_count);
  apportion(c->N, c->R, c->n_start, c->n_count);
  c->F_local = c->f_count[c->rank];
  c->f0 = c->f_start[c->rank];
  c->N_local = c->n_count[c->rank];
  c->n0 = c->n_start[c->rank];
  c->capacity = cfg->frame_capacity > 0 ? std::min(cfg->frame_capacity, std::max(c->F_local, 1))
                                        : std::max(c->F_local, 1);
  c->batch = cfg->batch_frames > 0 ? cfg->batch_frames : 256;
  c->batch = std::min(c->batch, c->capacity);
  c->cams.resize(cfg->n_cams);
  *out = c;
  int rc = [&]() -> int {
    CU(cudaSetDevice(cfg->device));
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    {
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU(cudaStreamCreateWithPriority(&c->stream_b, cudaStreamNonBlocking, hi));
      for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&c->ev_front[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_dec[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_back[i], cudaEventDisableTiming));
      }
      CU(cudaEventCreateWithFlags(&c->ev_tabs, cudaEventDisableTiming));
      for (int j = 0; j < upsp_gpu_ctx::NX; ++j) {
        CU(cudaStreamCreateWithFlags(&c->stream_x[j], cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&c->ev_x[i][j], cudaEventDisableTiming));
      }
      CU(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
    }
    CU(cudaEventCreateWithFlags(&c->ev_push, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_proc, cudaEventDisableTiming));
    CU(cudaEventCreate(&c->ev_a));
    CU(cudaEventCreate(&c->ev_b));
    CU(cudaEventCreate(&c->ev_pa));
    CU(cudaEventCreate(&c->ev_pb));
    CU(cudaEventCreate(&c->ev_t0));
    CU(cudaEventCreate(&c->ev_t1));
    c->proc_recs.resize(16);
    for (auto& r : c->proc_recs) {
      r.off = r.count = 0;
      CU(cudaEventCreateWithFlags(&r.ev, cudaEventDisableTiming));
    }
    const size_t nf = (size_t)c->N_local * c->F;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    c->off_sum = al(nf * sizeof(float));
    c->off_sumsq = c->off_sum + al((size_t)c->N * sizeof(double));
    c->shared_bytes = c->off_sumsq + al((size_t)c->N * sizeof(double));
    // UPSP_FORCE_VMM=1: the peer-mappable (CUDA VMM) allocation on a single rank too (A/B: is the allocation kind what
    // multi-rank phase 1 pays for?)
    if (c->R > 1 || (getenv("UPSP_FORCE_VMM") && atoi(getenv("UPSP_FORCE_VMM")))) {
      CU(cudaFree(0));  // make sure the primary context exists before driver-API calls
      TRY(vmm_alloc(c->shared_vmm, c->shared_bytes, cfg->device));
      c->d_shared = reinterpret_cast<char*>(c->shared_vmm.ptr);
    } else {
      TRY(dmalloc(&c->d_shared, c->shared_bytes));
    }
    c->d_itrans = reinterpret_cast<float*>(c->d_shared);
    c->d_sum = reinterpret_cast<double*>(c->d_shared + c->off_sum);
    c->d_sumsq = reinterpret_cast<double*>(c->d_shared + c->off_sumsq);
    CU(cudaMemsetAsync(c->d_sum, 0, (size_t)c->N * sizeof(double), c->stream));
    CU(cudaMemsetAsync(c->d_sumsq, 0, (size_t)c->N * sizeof(double), c->stream));
    TRY(dmalloc(&c->d_avg, c->N));
    TRY(dmalloc(&c->d_rms, c->N));
    TRY(dmalloc(&c->d_cov, c->N));
    TRY(dmalloc(&c->d_steady, c->N));
    TRY(dmalloc(&c->d_temp, c->N));
    TRY(dmalloc(&c->d_rms2, c->N_local));
    TRY(dmalloc(&c->d_avg2, c->N_local));
    TRY(dmalloc(&c->d_gain2, c->N_local));
    TRY(dmalloc(&c->d_rms2f, c->N_local));
    TRY(dmalloc(&c->d_avg2f, c->N_local));
    TRY(dmalloc(&c->d_gain2f, c->N_local));
    if (c->R > 1) {
      TRY(dmalloc(&c->d_tsum, c->N));
      TRY(dmalloc(&c->d_tsq, c->N));
      TRY(dmalloc(&c->d_peer_ptrs, 2 * c->R));
    }
    c->peer_base[c->rank] = c->d_shared;
    if (c->R == 1) c->peers_ready = true;
    CU(cudaStreamSynchronize(c->stream));
    return UPSP_OK;
  }();
  if (rc) {
    std::string keep = g_err;
    upsp_gpu_destroy(c);
    g_err = keep;
    *out = nullptr;
  }
  return rc;
}

static void free_camera(Camera& cam) {
  cudaFree(cam.d_in);
  cudaFree(cam.d_work);
  cudaFree(cam.d_work2);
  cudaFree(cam.d_hot_cnt2);
  cudaFree(cam.d_hot_pos2);
  cudaFree(cam.d_pv2);
  cudaFree(cam.d_warp);
  cudaFree(cam.d_hot_cnt);
  cudaFree(cam.d_hot_pos);
  cudaFree(cam.d_m6);
  cudaFree(cam.d_rho);
  cudaFree(cam.d_iters);
  cudaFree(cam.d_tab);
  cudaFree(cam.d_coef);
  cudaFree(cam.d_skip);
  cudaFree(cam.d_ref16);
  cudaFree(cam.d_eccT);
  cudaFree(cam.d_eccI);
  cudaFree(cam.d_eccTmp);
  cudaFree(cam.d_eccG);
  cudaFree(cam.d_eccPart);
  cudaFree(cam.d_eccState);
  cudaFree(cam.d_eccTab);
  cudaFree(cam.d_nactive);
  for (void* p : cam.patch_allocs) cudaFree(p);
  cudaFree(cam.d_cl_list);
  cudaFree(cam.d_scratch);
  cudaFree(cam.d_pv);
  cudaFree(cam.d_img32);
  cudaFree(cam.d_img32b);
  cudaFree(cam.d_filt16);
  cudaFree(cam.d_slot_pix);
  cudaFree(cam.d_code);
  cudaFree(cam.d_fix[0]);
  cudaFree(cam.d_fix[1]);
  cudaFree(cam.d_val);
  cudaFree(cam.d_rowptr);
}

extern "C" int upsp_gpu_destroy(upsp_gpu_ctx* c) {
  if (!c) return UPSP_OK;
  cudaSetDevice(c->cfg.device);
  if (c->stream_b) cudaStreamSynchronize(c->stream_b);
  for (auto sx : c->stream_x) if (sx) cudaStreamSynchronize(sx);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
  for (int r = 0; r < c->R; ++r)
    if (r != c->rank && c->peer_is_ipc[r]) vmm_free(c->peer_vmm[r]);
  for (auto& cam : c->cams) free_camera(cam);
  cudaFree(c->d_lut);
  cudaFree(c->d_perm);
  cudaFree(c->d_perm_tma);
  cudaFree(c->d_other_idx);
  cudaFree(c->d_other_local);
  cudaFree(c->d_bounce);
  cudaFree(c->d_p2coef);
  cudaFree(c->d_cl_parts);
  cudaFree(c->d_p2parts);
  cudaFree(c->d_tma_blk);
  cudaFree(c->d_intensity);
  if (c->shared_vmm.handle) vmm_free(c->shared_vmm); else cudaFree(c->d_shared);
  if (c->ptrans_owned) cudaFree(c->d_ptrans);
  cudaFree(c->d_avg);
  cudaFree(c->d_rms);
  cudaFree(c->d_cov);
  cudaFree(c->d_steady);
  cudaFree(c->d_temp);
  cudaFree(c->d_rms2);
  cudaFree(c->d_avg2);
  cudaFree(c->d_gain2);
  cudaFree(c->d_rms2f);
  cudaFree(c->d_avg2f);
  cudaFree(c->d_gain2f);
  cudaFree(c->d_tsum);
  cudaFree(c->d_tsq);
  cudaFree(c->d_peer_ptrs);
  cudaFree(c->d_nccl_send);
  cudaFree(c->d_nccl_recv);
  if (c->nccl_comm && nccl().ok) nccl().CommDestroy(c->nccl_comm);
  if (c->ev_push) cudaEventDestroy(c->ev_push);
  if (c->ev_proc) cudaEventDestroy(c->ev_proc);
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  if (c->ev_pa) cudaEventDestroy(c->ev_pa);
  if (c->ev_t0) cudaEventDestroy(c->ev_t0);
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  for (auto& r : c->proc_recs) if (r.ev) cudaEventDestroy(r.ev);
  for (auto e : c->kev) cudaEventDestroy(e);
  if (c->ev_pb) cudaEventDestroy(c->ev_pb);
  for (int i = 0; i < 2; ++i) {
    if (c->ev_front[i]) cudaEventDestroy(c->ev_front[i]);
    if (c->ev_dec[i]) cudaEventDestroy(c->ev_dec[i]);
    if (c->ev_back[i]) cudaEventDestroy(c->ev_back[i]);
  }
  if (c->ev_tabs) cudaEventDestroy(c->ev_tabs);
  for (int i = 0; i < 2; ++i) {
    for (auto e : c->ev_x[i]) if (e) cudaEventDestroy(e);
    cudaFree(c->d_stage[i]);
    c->d_stage[i] = nullptr;
  }
  for (auto sx : c->stream_x) if (sx) cudaStreamDestroy(sx);
  if (c->stream_b) cudaStreamDestroy(c->stream_b);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  delete c;
  return UPSP_OK;
}

extern "C" int upsp_gpu_get_slices(const upsp_gpu_ctx* c, int* ff, int* nf, int* fn, int* nn) {
  REQUIRE(c, UPSP_ERR_INVALID, "null context");
  if (ff) *ff = c->f0;
  if (nf) *nf = c->F_local;
  if (fn) *fn = c->n0;
  if (nn) *nn = c->N_local;
  return UPSP_OK;
}

// ------------------------------------------------------------------------------------------
// setup
// ------------------------------------------------------------------------------------------
#define CAM_CHECK(c, cam) \
  REQUIRE((cam) >= 0 && (cam) < (int)(c)->cams.size(), UPSP_ERR_INVALID, "camera %d out of range", cam)
#define NOT_FINAL(c) \
  REQUIRE(!(c)->finalized, UPSP_ERR_STATE, "setup calls are not allowed after the first process_frames")

extern "C" int upsp_gpu_set_camera(upsp_gpu_ctx* c, int cam, int width, int height) {
  ENTER(c);
  CAM_CHECK(c, cam);
  NOT_FINAL(c);
  REQUIRE(width > 0 && height > 0, UPSP_ERR_INVALID, "bad frame size %dx%d", width, height);
  c->cams[cam].W = width;
  c->cams[cam].H = height;
  c->cams[cam].npix = (size_t)width * height;
  return UPSP_OK;
}

extern "C" int upsp_gpu_set_projection(upsp_gpu_ctx* c, int cam, const int32_t* rowptr,
                                       const int32_t* col, const float* val) {
  ENTER(c);
  CAM_CHECK(c, cam);
  NOT_FINAL(c);
  Camera& k = c->cams[cam];
  REQUIRE(k.npix > 0, UPSP_ERR_STATE, "set_camera(%d) first", cam);
  REQUIRE(rowptr, UPSP_ERR_INVALID, "null rowptr");
  REQUIRE(rowptr[0] == 0, UPSP_ERR_INVALID, "rowptr[0] != 0");
  for (int i = 0; i < c->N; ++i)
    REQUIRE(rowptr[i + 1] >= rowptr[i], UPSP_ERR_INVALID, "rowptr not monotone at row %d", i);
  const int nnz = rowptr[c->N];
  REQUIRE(nnz == 0 || (col && val), UPSP_ERR_INVALID, "null col/val");
  for (int i = 0; i < nnz; ++i)
    REQUIRE(col[i] >= 0 && (size_t)col[i] < k.npix, UPSP_ERR_INVALID,
            "column %d of entry %d outside the %dx%d frame", col[i], i, k.W, k.H);
  k.rowptr.assign(rowptr, rowptr + c->N + 1);
  k.col.assign(col, col + nnz);
  k.val.assign(val, val + nnz);
  k.has_proj = true;
  return UPSP_OK;
}

extern "C" int upsp_gpu_set_overlap_remap(upsp_gpu_ctx* c, const int32_t* src) {
  ENTER(c);
  NOT_FINAL(c);
  if (!src) {
    c->remap.clear();
    return UPSP_OK;
  }
  for (int i = 0; i < c->N; ++i)
    REQUIRE(src[i] >= 0 && src[i] < c->N, UPSP_ERR_INVALID, "src_index[%d]=%d out of range", i, src[i]);
  c->remap.assign(src, src + c->N);
  return UPSP_OK;
}

extern "C" int upsp_gpu_set_options(upsp_gpu_ctx* c, int registration, int interp, int patcher,
                                    int hot_pixel_fix) {
  ENTER(c);
  NOT_FINAL(c);
  REQUIRE(registration >= UPSP_REG_NONE && registration <= UPSP_REG_GIVEN, UPSP_ERR_INVALID,
          "registration %d", registration);
  REQUIRE(interp == UPSP_INTERP_NEAREST || interp == UPSP_INTERP_LINEAR, UPSP_ERR_INVALID,
          "interp %d", interp);
  REQUIRE(patcher == UPSP_PATCH_NONE || patcher == UPSP_PATCH_POLYNOMIAL, UPSP_ERR_INVALID,
          "patcher %d", patcher);
  c->registration = registration;
  c->interp = interp;
  c->patcher = patcher;
  c->hot_fix = hot_pixel_fix != 0;
  return UPSP_OK;
}

#include "gauss_fixed.inc"

extern "C" int upsp_gpu_set_filter(upsp_gpu_ctx* c, int kind, int ksize) {
  ENTER(c);
  NOT_FINAL(c);
  REQUIRE(kind >= 0 && kind <= 2, UPSP_ERR_INVALID, "filter kind %d", kind);
  c->filter = FilterSpec{};
  if (kind == 0) return UPSP_OK;
  REQUIRE(ksize >= 1 && (ksize & 1), UPSP_ERR_INVALID, "filter_size must be odd (psp_process.cpp:1296), got %d", ksize);
  if (ksize == 1) return UPSP_OK;   // a 1x1 Gaussian / box window is the identity
  if (kind == 1) {
    REQUIRE(ksize >= 3 && ksize <= 31, UPSP_ERR_INVALID, "gaussian filter_size %d is not built (odd sizes 3..31)", ksize);
    const int* half = kGaussFixed[(ksize - 3) / 2];
    const int r = ksize / 2;
    for (int i = 0; i <= r; ++i) c->filter.kq[i] = c->filter.kq[ksize - 1 - i] = half[i];
    if (ksize <= 7)   // dyadic taps: the float kernel of the CV_32F path is the same numbers
      for (int i = 0; i < ksize; ++i) c->filter.kf[i] = (float)((double)c->filter.kq[i] / 65536.0);
  } else {
    REQUIRE(ksize <= 31, UPSP_ERR_INVALID, "box filter_size %d too large", ksize);
  }
  c->filter.kind = kind;
  c->filter.ksize = ksize;
  return UPSP_OK;
}

extern "C" int upsp_gpu_set_unpack_lut(upsp_gpu_ctx* c, const uint16_t* lut) {
  ENTER(c);
  // one table per context, and its largest value decides which projection kernels are eligible (12-bit fast paths):
  // it cannot change once frames have been decoded with it
  REQUIRE(c->frames_processed == 0 && !c->finalized, UPSP_ERR_STATE,
          "set_unpack_lut after frames were processed (the table is fixed for the run)");
  cudaFree(c->d_lut);
  c->d_lut = nullptr;
  c->lut_max = 0;
  if (lut) {
    TRY(upload(&c->d_lut, lut, 1024));
    for (int i = 0; i < 1024; ++i) c->lut_max = std::max(c->lut_max, (int)lut[i]);
  }
  return UPSP_OK;
}

extern "C" int upsp_gpu_set_reference_frame(upsp_gpu_ctx* c, int cam, const uint16_t* frame) {
  ENTER(c);
  CAM_CHECK(c, cam);
  Camera& k = c->cams[cam];
  REQUIRE(k.npix > 0, UPSP_ERR_STATE, "set_camera(%d) first", cam);
  REQUIRE(frame, UPSP_ERR_INVALID, "null frame");
  cudaFree(k.d_ref16);
  k.d_ref16 = nullptr;
  TRY(upload(&k.d_ref16, frame, k.npix));
  k.has_ref = true;
  return UPSP_OK;
}

extern "C" int upsp_gpu_set_warp_matrices(upsp_gpu_ctx* c, int cam, int off, int count,
                                          const float* m6) {
  ENTER(c);
  CAM_CHECK(c, cam);
  REQUIRE(off >= 0 && count >= 0 && off + count <= c->F_local, UPSP_ERR_INVALID,
          "frames [%d,%d) outside the local slice of %d", off, off + count, c->F_local);
  REQUIRE(m6 || count == 0, UPSP_ERR_INVALID, "null matrices");
  Camera& k = c->cams[cam];
  if (!k.d_m6) {
    TRY(dmalloc(&k.d_m6, (size_t)std::max(c->F_local, 1) * 6));
    std::vector<float> id((size_t)std::max(c->F_local, 1) * 6, 0.0f);
    for (int f = 0; f < c->F_local; ++f) id[(size_t)f * 6 + 0] = id[(size_t)f * 6 + 4] = 1.0f;
    CU(cudaMemcpy(k.d_m6, id.data(), id.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (count)
    CU(cudaMemcpyAsync(k.d_m6 + (size_t)off * 6, m6, (size_t)count * 6 * sizeof(float),
                       cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  k.has_m6 = true;
  return UPSP_OK;
}

extern "C" int upsp_gpu_set_patches(upsp_gpu_ctx* c, int cam, int ncl, const int32_t* boff,
                                    const uint32_t* bx, const uint32_t* by, const int32_t* ioff,
                                    const uint32_t* ix, const uint32_t* iy) {
  ENTER(c);
  CAM_CHECK(c, cam);
  NOT_FINAL(c);
  Camera& k = c->cams[cam];
  REQUIRE(k.npix > 0, UPSP_ERR_STATE, "set_camera(%d) first", cam);
  REQUIRE(!k.has_patches, UPSP_ERR_STATE, "patches of camera %d already set", cam);
  REQUIRE(ncl >= 0, UPSP_ERR_INVALID, "n_clusters %d", ncl);
  if (ncl == 0) return UPSP_OK;
  REQUIRE(boff && ioff, UPSP_ERR_INVALID, "null offsets");
  const int nbt = boff[ncl], nit = ioff[ncl];
  REQUIRE(boff[0] == 0 && ioff[0] == 0, UPSP_ERR_INVALID, "offsets must start at 0");
  for (int i = 0; i < nbt; ++i)
    REQUIRE((int)bx[i] < k.W && (int)by[i] < k.H, UPSP_ERR_INVALID, "boundary pixel outside frame");
  for (int i = 0; i < nit; ++i)
    REQUIRE((int)ix[i] < k.W && (int)iy[i] < k.H, UPSP_ERR_INVALID, "interior pixel outside frame");

  std::vector<int> bsrc(nbt), nzp(ncl), perm((size_t)ncl * 10, 0), islot_pix(nit), level(ncl, 0);
  std::vector<float> qr_e((size_t)10 * nbt, 0.0f), qr_te((size_t)10 * nbt, 0.0f),
      hcoef((size_t)ncl * 10, 0.0f), ipow((size_t)6 * nit);
  std::unordered_map<int, std::pair<int, int>> owner;  // pixel -> (cluster, slot) of earlier active clusters
  int max_level = 0;
  for (int cl = 0; cl < ncl; ++cl) {
    const int o = boff[cl], nb = boff[cl + 1] - o;
    REQUIRE(nb >= 0 && ioff[cl + 1] >= ioff[cl], UPSP_ERR_INVALID, "offsets not monotone");
    nzp[cl] = -1;                    // inactive: fewer than (3+2)(3+1)/2 = 10 boundary px
    if (nb < 10) continue;           // patches.ipp:112-115
    std::vector<float> A((size_t)nb * 10);
    for (int i = 0; i < nb; ++i) {
      const int pix = (int)by[o + i] * k.W + (int)bx[o + i];
      auto it = owner.find(pix);
      if (it != owner.end()) {
        bsrc[o + i] = -1 - it->second.second;
        level[cl] = std::max(level[cl], level[it->second.first] + 1);
      } else {
        bsrc[o + i] = pix;
      }
      int cnt = 0;  // patches.ipp:183-195: A(ind,count) = (T)pow(y,i) * (T)pow(x,j)
      for (int a = 0; a <= 3; ++a)
        for (int b = 0; b <= 3; ++b)
          if (a + b <= 3) {
            A[(size_t)cnt * nb + i] =
                (float)std::pow((double)by[o + i], a) * (float)std::pow((double)bx[o + i], b);
            ++cnt;
          }
    }
    ColPivQR f = colpiv_householder_qr(std::move(A), nb, 10);
    nzp[cl] = f.nonzero_pivots;
    for (int q = 0; q < 10; ++q) {
      hcoef[(size_t)cl * 10 + q] = f.hcoef[q];
      perm[(size_t)cl * 10 + q] = f.perm[q];
      for (int i = 0; i < nb; ++i) {
        const float e = f.qr[(size_t)q * nb + i];
        qr_e[(size_t)10 * o + (size_t)q * nb + i] = e;
        qr_te[(size_t)10 * o + (size_t)q * nb + i] = f.hcoef[q] * e;
      }
    }
    max_level = std::max(max_level, level[cl]);
    k.max_bounds = std::max(k.max_bounds, nb);
    for (int j = ioff[cl]; j < ioff[cl + 1]; ++j) {
      const int pix = (int)iy[j] * k.W + (int)ix[j];
      owner[pix] = {cl, j};
      k.pix2slot[pix] = j;
    }
  }
  for (int j = 0; j < nit; ++j) {
    islot_pix[j] = (int)iy[j] * k.W + (int)ix[j];
    for (int p = 1; p <= 3; ++p) {
      ipow[(size_t)6 * j + p - 1] = (float)std::pow((double)ix[j], p);
      ipow[(size_t)6 * j + 2 + p] = (float)std::pow((double)iy[j], p);
    }
  }
  k.levels.assign(max_level + 1, {});
  for (int cl = 0; cl < ncl; ++cl)
    if (nzp[cl] >= 0) k.levels[level[cl]].push_back(cl);
  std::vector<int> cl_list;
  k.level_off.assign(1, 0);
  for (auto& l : k.levels) {
    cl_list.insert(cl_list.end(), l.begin(), l.end());
    k.level_off.push_back((int)cl_list.size());
  }
  auto up = [&](auto** dptr, const auto& h) -> int {
    typedef typename std::remove_reference<decltype(h[0])>::type T;
    typename std::remove_const<T>::type* d = nullptr;
    TRY(upload(&d, h.data(), h.size()));
    *dptr = d;
    k.patch_allocs.push_back((void*)d);
    return UPSP_OK;
  };
  std::vector<int> boff_v(boff, boff + ncl + 1), ioff_v(ioff, ioff + ncl + 1);
  PatchGeom& g = k.geom;
  g.n_clusters = ncl;
  TRY(up(&g.bounds_off, boff_v));
  TRY(up(&g.bsrc, bsrc));
  TRY(up(&g.qr_e, qr_e));
  TRY(up(&g.qr_te, qr_te));
  TRY(up(&g.hcoef, hcoef));
  TRY(up(&g.perm, perm));
  TRY(up(&g.nzp, nzp));
  TRY(up(&g.internal_off, ioff_v));
  TRY(up(&g.ipow, ipow));
  TRY(up(&g.islot_pix, islot_pix));
  TRY(upload(&k.d_cl_list, cl_list.data(), cl_list.size()));
  k.total_bounds = nbt;
  k.total_internal = nit;
  k.has_patches = true;
  return UPSP_OK;
}

// Fold remap + patched-pixel codes into device tables; allocate per-batch working buffers.
static int finalize(upsp_gpu_ctx* c) {
  if (c->finalized) return UPSP_OK;
  const int N = c->N;
  for (size_t i = 0; i < c->cams.size(); ++i) {
    REQUIRE(c->cams[i].npix > 0, UPSP_ERR_STATE, "camera %zu has no frame size", i);
    REQUIRE(c->cams[i].has_proj, UPSP_ERR_STATE, "camera %zu has no projection matrix", i);
  }
  const bool use_patch = c->patcher == UPSP_PATCH_POLYNOMIAL;
  // is every remapped row <= 1 entry in every camera?
  c->ell1 = true;
  for (auto& k : c->cams)
    for (int n = 0; n < N && c->ell1; ++n)
      if (k.rowptr[n + 1] - k.rowptr[n] > 1) c->ell1 = false;
  // the NCCL exchange needs the frame-major intermediate (local transpose into send blocks + ncclSend/ncclRecv)
  c->fused = c->ell1 && c->cfg.keep_frame_major == 0 && c->filter.kind == 0 && c->exchange == UPSP_XCHG_PEER;
  // two-stream pipeline: front end (decode, hot pixels, patch) of batch i+1 overlaps the fused
  // projection of batch i.  Not with the device ECC solve (its blur/moment kernels want the whole
  // GPU) and not in the unfused modes.  UPSP_PIPELINE=0 turns it off.
  c->pipelined = c->fused && c->registration != UPSP_REG_PIXEL &&
                 !(getenv("UPSP_PIPELINE") && atoi(getenv("UPSP_PIPELINE")) == 0);
  std::vector<float> cov(N, 0.0f);
  for (size_t ci = 0; ci < c->cams.size(); ++ci) {
    Camera& k = c->cams[ci];
    auto code_of = [&](int col) -> int {
      if (use_patch && k.has_patches && c->filter.kind == 0) {   // filtered images carry the patch in-pixel
        auto it = k.pix2slot.find(col);
        if (it != k.pix2slot.end()) return -2 - it->second;
      }
      return col;
    };
    std::vector<int> code, rp;
    std::vector<float> val;
    if (c->ell1) {
      code.assign(N, -1);
      val.assign(N, 0.0f);
      for (int n = 0; n < N; ++n) {
        const int s = c->remap.empty() ? n : c->remap[n];
        if (k.rowptr[s + 1] > k.rowptr[s]) {
          code[n] = code_of(k.col[k.rowptr[s]]);
          val[n] = k.val[k.rowptr[s]];
        }
      }
    } else {
      rp.assign(N + 1, 0);
      for (int n = 0; n < N; ++n) {
        const int s = c->remap.empty() ? n : c->remap[n];
        rp[n + 1] = rp[n] + (k.rowptr[s + 1] - k.rowptr[s]);
      }
      code.resize(rp[N]);
      val.resize(rp[N]);
      for (int n = 0; n < N; ++n) {
        const int s = c->remap.empty() ? n : c->remap[n];
        for (int j = 0; j < k.rowptr[s + 1] - k.rowptr[s]; ++j) {
          code[rp[n] + j] = code_of(k.col[k.rowptr[s] + j]);
          val[rp[n] + j] = k.val[k.rowptr[s] + j];
        }
      }
      TRY(upload(&k.d_rowptr, rp.data(), rp.size()));
    }
    TRY(upload(&k.d_code, code.data(), code.size()));
    TRY(upload(&k.d_val, val.data(), val.size()));
    if (c->ell1) {
      k.h_code = code;
      k.h_val = val;
    }
    // coverage = sum_c project(ones) (psp_process.cpp:1953-1966), then adjust_solution (:1974)
    for (int n = 0; n < N; ++n) {
      const int s = c->remap.empty() ? n : c->remap[n];
      float t = 0.0f;
      for (int j = k.rowptr[s]; j < k.rowptr[s + 1]; ++j) t += k.val[j] * 1.0f;
      t = 0.0f + t;
      cov[n] = ci == 0 ? t : cov[n] + t;
    }
    // working buffers
    TRY(dmalloc(&k.d_work, (size_t)c->batch * k.npix));
    TRY(dmalloc(&k.d_hot_cnt, (size_t)2 * c->batch));   // [hot counts | finished-block tickets]
    TRY(dmalloc(&k.d_hot_pos, (size_t)c->batch * UPSP_HOT_STORE));
    if (c->pipelined) {
      TRY(dmalloc(&k.d_work2, (size_t)c->batch * k.npix));
      TRY(dmalloc(&k.d_hot_cnt2, (size_t)2 * c->batch));
      TRY(dmalloc(&k.d_hot_pos2, (size_t)c->batch * UPSP_HOT_STORE));
    }
    if (c->registration != UPSP_REG_NONE) {
      if (!c->fused) TRY(dmalloc(&k.d_warp, (size_t)c->batch * k.npix));
      TRY(dmalloc(&k.d_tab, (size_t)std::max(c->F_local, 1) * (2 * k.W + 2 * k.H)));
      TRY(dmalloc(&k.d_coef, (size_t)std::max(c->F_local, 1)));
      if (c->registration == UPSP_REG_GIVEN)
        REQUIRE(k.has_m6, UPSP_ERR_STATE, "registration=given but camera %zu has no warp matrices", ci);
      if (c->registration == UPSP_REG_PIXEL) {
        REQUIRE(k.has_ref, UPSP_ERR_STATE,
                "registration=pixel needs the first frame of camera %zu (upsp_gpu_set_reference_frame)", ci);
        const int eb = std::min(c->batch, ECC_BATCH);
        if (!k.d_m6) TRY(dmalloc(&k.d_m6, (size_t)std::max(c->F_local, 1) * 6));
        TRY(dmalloc(&k.d_eccT, k.npix));
        TRY(dmalloc(&k.d_eccI, (size_t)eb * k.npix));
        TRY(dmalloc(&k.d_eccTmp, (size_t)eb * k.npix));
        TRY(dmalloc(&k.d_eccG, (size_t)eb * k.npix));
        TRY(dmalloc(&k.d_eccPart, (size_t)eb * ECC_NBLK * ECC_NSUM));
        TRY(dmalloc(&k.d_eccState, (size_t)std::max(c->F_local, 1)));
        TRY(dmalloc(&k.d_eccTab, (size_t)eb * (2 * k.W + 2 * k.H)));
        TRY(dmalloc(&k.d_nactive, 1));
        // template = GaussianBlur(first frame as f32)  (ecc.cpp: templateFloat)
        const dim3 g(cdiv(k.W, 256), k.H, 1);
        k_ecc_blur_rows<uint16_t><<<g, 256, 0, c->stream>>>(k.d_ref16, k.d_eccTmp, k.W, k.H);
        KCHECK(c);
        k_ecc_blur_cols<<<g, 256, 0, c->stream>>>(k.d_eccTmp, k.d_eccT, k.W, k.H);
        KCHECK(c);
      }
    }
    if (use_patch && k.has_patches) {
      TRY(dmalloc(&k.d_pv, (size_t)std::max(k.total_internal, 1) * c->batch));
      if (c->pipelined) TRY(dmalloc(&k.d_pv2, (size_t)std::max(k.total_internal, 1) * c->batch));
    }
    if (c->filter.kind) {
      if (use_patch && k.has_patches) {
        REQUIRE(c->filter.kind != 1 || c->filter.ksize <= 7, UPSP_ERR_INVALID,
                "gaussian filter_size %d on the patched (CV_32F) image is not built: 3, 5, 7 only", c->filter.ksize);
        TRY(dmalloc(&k.d_img32, (size_t)c->batch * k.npix));
        TRY(dmalloc(&k.d_img32b, (size_t)c->batch * k.npix));
        std::vector<int> slot_pix(std::max(k.total_internal, 1), -1);
        for (auto& kv : k.pix2slot) slot_pix[kv.second] = kv.first;    // final writer of each pixel
        TRY(upload(&k.d_slot_pix, slot_pix.data(), slot_pix.size()));
      } else {
        TRY(dmalloc(&k.d_filt16, (size_t)c->batch * k.npix));
      }
    }
  }
  CU(cudaMemcpy(c->d_cov, cov.data(), (size_t)N * sizeof(float), cudaMemcpyHostToDevice));
  if (c->fused) {
    // processing order: raster order of each node's pixel in the first camera that sees it, so
    // the 32 nodes of a warp read (mostly) one image row: a warp-wide tap load then touches 1-2
    // 128-byte lines instead of ~7 for a square (Z-order) patch or ~16 for mesh order -- the
    // kernel is bound by L1 wavefronts, not bytes.  Nodes without a plain pixel (skipped /
    // patched-only) go last.  Pure locality hint: results do not depend on it.
    std::vector<std::pair<uint64_t, int>> keyed(N);
    for (int n = 0; n < N; ++n) {
      uint64_t key = ~0ull;
      const int sn = c->remap.empty() ? n : c->remap[n];
      for (size_t ci = 0; ci < c->cams.size(); ++ci) {
        const Camera& k = c->cams[ci];
        if (k.rowptr[sn + 1] > k.rowptr[sn]) {
          const int col = k.col[k.rowptr[sn]];
          key = ((uint64_t)ci << 32) | (uint32_t)col;
          break;
        }
      }
      keyed[n] = {key, n};
    }
    std::sort(keyed.begin(), keyed.end());
    std::vector<int> perm(N);
    for (int i = 0; i < N; ++i) perm[i] = keyed[i].second;
    TRY(upload(&c->d_perm, perm.data(), perm.size()));
  }
  // Staged exchange or direct peer stores from the projection kernel?  Measured on 8 x B200 (NVSwitch):
  // one GPU pair moves ~300 GB/s whichever way it is driven, and the copy engines move ~290 GB/s per
  // GPU in total, while the SMs' direct stores fan out over all pairs (~570 GB/s per GPU at 8 ranks)
  // but make the projection kernel wait on NVLink.
  //   2 ranks: copy engines win (phase 1 of 20k frames/GPU: 82 ms direct, 56 ms staged);
  //   8 ranks: direct stores in 128-byte segments win (78.7 ms; all staged 121 ms; mixed 3 staged +
  //            4 direct 86 ms, 2 + 5: 87 ms).
  // So: `staged_peers` = 1 at 2 ranks, 0 otherwise; UPSP_STAGED_PEERS=k routes the next k ranks' rows
  // through the staging block (0: all direct).
  // Round 2: the staged rows can also be shipped by a small SM kernel (k_ship_rows: 1 KB row pieces per batch, all
  // peers interleaved) instead of the copy engines: UPSP_SHIP = sm | ce, UPSP_SHIP_BPSM = blocks of 128 threads per SM.
  c->ship_sm = getenv("UPSP_SHIP") && !strcmp(getenv("UPSP_SHIP"), "sm");
  c->ship_bpsm = getenv("UPSP_SHIP_BPSM") ? std::max(1, atoi(getenv("UPSP_SHIP_BPSM"))) : 1;
  c->staged_peers = c->R == 2 ? 1 : 0;
  if (getenv("UPSP_STAGED_PEERS")) c->staged_peers = std::min(std::max(atoi(getenv("UPSP_STAGED_PEERS")), 0), c->R - 1);
  if (!c->pipelined || c->R <= 1) c->staged_peers = 0;
  c->staged_xchg = c->staged_peers > 0;
  c->stage_mask = 0;
  for (int d = 1; d <= c->staged_peers; ++d) c->stage_mask |= 1u << ((c->rank + d) % c->R);
  if (c->staged_xchg) {
    c->stage_stride = (c->batch + 3) & ~3;
    for (int i = 0; i < 2; ++i) TRY(dmalloc(&c->d_stage[i], (size_t)c->N * c->stage_stride));
  }
  // big buffers that depend on the mode
  {
    const size_t fn = (size_t)c->F_local * c->N, nf = (size_t)c->N_local * c->F;
    const bool alias = c->cfg.pressure_aliases_intensity != 0;
    if (!c->fused) {
      TRY(dmalloc(&c->d_intensity, alias ? std::max(fn, nf) : fn));
      if (alias) c->d_ptrans = c->d_intensity;
    }
    if (!c->d_ptrans) {
      TRY(dmalloc(&c->d_ptrans, nf));
      c->ptrans_owned = true;
    }
  }
  c->finalized = true;
  return UPSP_OK;
}


// column coefficients of the warp maps as the TMA projection reads them from constant memory: (M0, M3) * 1024 in
// double (k_warp_tables' own expression), identity for the frame that is never registered (psp_process.cpp:1777)
__global__ void k_tma_coef(const float* __restrict__ m6, int n, int skip, double2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = i == skip ? make_double2(1024.0, 0.0) : make_double2((double)m6[(size_t)i * 6] * 1024.0, (double)m6[(size_t)i * 6 + 3] * 1024.0);
}

// ------------------------------------------------------------------------------------------
// TMA-staged projection: block partition + tensor maps (kernels_project_tma.cuh)
// ------------------------------------------------------------------------------------------
static int encode_tmap(CUtensorMap* m, CUtensorMapDataType dt, void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2 = 1) {
  DriverApi& d = drv();
  REQUIRE(d.TensorMapEncodeTiled != nullptr, UPSP_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  const cuuint32_t box[3] = {b0, b1, b2};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = d.TensorMapEncodeTiled(m, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REQUIRE(r == CUDA_SUCCESS, UPSP_ERR_CUDA, "cuTensorMapEncodeTiled -> %d", (int)r);
  return UPSP_OK;
}

static int phase2_cluster(int F, bool in16 = false);

// Decide the projection mode of a fused single-camera context and build what the TMA kernels need:
//   * processing order: nodes with a plain pixel, cut into blocks of <= 128 nodes whose pixels lie in one
//     strip of `strip_rows` image rows and span <= `tile_cols` columns (so that the box a block stages per
//     frame has a fixed size), raster order inside a block; then the nodes without a plain pixel
//     (patched / unseen), which k_project_fused4 handles;
//   * tensor maps over the decoded work buffers (mode 1) or the packed input store (mode 2).
// UPSP_PROJ = v4 | tma16 | tma12 overrides the choice (A/B measurements and the parity test that pins the
// three to identical bits).
static int ensure_proj_mode(upsp_gpu_ctx* c) {
  if (c->proj_mode >= 0) return UPSP_OK;
  c->proj_mode = 0;
  if (!c->fused || c->cams.size() != 1 || c->registration == UPSP_REG_NONE || c->interp != UPSP_INTERP_LINEAR) return UPSP_OK;
  Camera& k = c->cams[0];
  const bool pix13 = k.format == UPSP_PIX_PACKED12 || (k.format == UPSP_PIX_PACKED10 && c->lut_max < 8192);
  if (!pix13 || (size_t)c->batch * k.npix >= ((size_t)1 << 31) || (size_t)c->batch * (size_t)(k.W + k.H) >= ((size_t)1 << 31))
    return UPSP_OK;
  if (drv().TensorMapEncodeTiled == nullptr || c->batch > tma_max_batch()) return UPSP_OK;
  const bool ok16 = k.W % 8 == 0;
  const bool ok12 = k.format == UPSP_PIX_PACKED12 && k.W % 32 == 0 && k.npix % 32 == 0 && k.frame_bytes % 16 == 0 &&
                    c->registration == UPSP_REG_GIVEN && c->pipelined;
  int want = ok12 ? 2 : (ok16 ? 1 : 0);
  if (getenv("UPSP_FUSED_V1") && atoi(getenv("UPSP_FUSED_V1"))) want = 0;
  if (const char* e = getenv("UPSP_PROJ")) {
    if (!strcmp(e, "v4")) want = 0;
    else if (!strcmp(e, "tma16")) want = ok16 ? 1 : 0;
    else if (!strcmp(e, "tma12")) want = ok12 ? 2 : (ok16 ? 1 : 0);
  }
  if (want == 0) return UPSP_OK;
  const TmaGeom g = tma_geom();
  const int N = c->N, W = k.W;
  std::vector<int> plain, other;
  for (int n = 0; n < N; ++n) (k.h_code[n] >= 0 ? plain : other).push_back(n);
  std::vector<std::pair<uint64_t, int>> keyed(plain.size());
  for (size_t i = 0; i < plain.size(); ++i) {
    const int code = k.h_code[plain[i]], y = code / W, x = code % W;
    keyed[i] = {((uint64_t)(y / g.strip_rows) << 42) | ((uint64_t)x << 21) | (uint64_t)y, plain[i]};
  }
  std::sort(keyed.begin(), keyed.end());
  std::vector<TmaBlock> blocks;
  std::vector<int> perm;
  const bool lane_assign = !(getenv("UPSP_TMA_LANES") && atoi(getenv("UPSP_TMA_LANES")) == 0);   // A/B knob
  perm.reserve(N);
  bool val1 = true;
  size_t i = 0;
  while (i < keyed.size()) {
    const int code0 = k.h_code[keyed[i].second];
    const int strip = (code0 / W) / g.strip_rows, x0 = code0 % W;
    size_t j = i;
    int ymin = 1 << 30, ymax = -1, xmax = x0;
    while (j < keyed.size() && (int)(j - i) < g.nodes_per_block) {
      const int code = k.h_code[keyed[j].second], y = code / W, x = code % W;
      if (y / g.strip_rows != strip || x - x0 > g.tile_cols) break;
      ymin = std::min(ymin, y);
      ymax = std::max(ymax, y);
      xmax = std::max(xmax, x);
      ++j;
    }
    std::vector<std::pair<int, int>> blk;      // (pixel, node): raster order inside the block
    for (size_t q = i; q < j; ++q) blk.push_back({k.h_code[keyed[q].second], keyed[q].second});
    std::sort(blk.begin(), blk.end());
    if (lane_assign) {
      // Which warp of the block takes which node: the four taps of a node-frame are shared-memory loads at
      // (row - box row) * row stride + (column - box column), the same offset for every node of a frame up to the
      // sub-pixel rotation, so WHICH banks the 32 lanes of a warp hit is fixed by the node pixels.  Nodes are dealt to
      // the block's warps so that as few lanes as possible share a bank without sharing the 32-bit word (raster order
      // put half a row and the start of the next one in a warp: 2.3 wavefronts per tap load, ncu r2i).
      const int nw = ((int)blk.size() + 31) / 32;
      std::vector<std::vector<std::pair<int, int>>> wn(nw);
      std::vector<std::array<std::array<std::vector<int>, 32>, 8>> words(nw);    // [warp][box phase][bank] -> distinct words
      const int nphase = want == 1 ? 2 : 8;      // column of the node inside the box modulo 2 px (u16) / 8 px (packed: 12 bytes)
      auto word_of = [&](int code, int par) {
        const int ry = code / W - ymin, rx = code % W - x0 + par;
        return want == 1 ? ry * (g.box_px16 / 2) + (rx >> 1) : ry * g.box_words12 + ((rx + (rx >> 1)) >> 2);   // u16 pair / packed byte floor(1.5 x)
      };
      for (auto& e : blk) {
        int best = -1, best_cost = 1 << 30;
        for (int w = 0; w < nw; ++w) {
          const int cap = std::min(32, (int)blk.size() - 32 * w);
          if ((int)wn[w].size() >= cap) continue;
          int cost = 0;
          for (int par = 0; par < nphase; ++par) {
            const int wd = word_of(e.first, par);
            const auto& v = words[w][par][wd & 31];
            if (!v.empty() && std::find(v.begin(), v.end(), wd) == v.end()) cost += (int)v.size();
          }
          if (cost < best_cost) {
            best_cost = cost;
            best = w;
          }
        }
        wn[best].push_back(e);
        for (int par = 0; par < nphase; ++par) {
          const int wd = word_of(e.first, par);
          auto& v = words[best][par][wd & 31];
          if (std::find(v.begin(), v.end(), wd) == v.end()) v.push_back(wd);
        }
      }
      blk.clear();
      for (auto& v : wn) blk.insert(blk.end(), v.begin(), v.end());
    }
    TmaBlock b{};
    b.node0 = (int)perm.size();
    b.count = (int)blk.size();
    b.xmin = x0;
    b.xmax = xmax;
    b.ymin = ymin;
    b.ymax = ymax;
    blocks.push_back(b);
    for (auto& e : blk) {
      perm.push_back(e.second);
      val1 = val1 && k.h_val[e.second] == 1.0f;
    }
    i = j;
  }
  c->n_tma_plain = (int)perm.size();
  c->n_tma_blocks = (int)blocks.size();
  c->tma_val1 = val1;
  perm.insert(perm.end(), other.begin(), other.end());
  TRY(upload(&c->d_perm_tma, perm.data(), perm.size()));
  TRY(upload(&c->d_tma_blk, blocks.data(), blocks.size()));
  if (want == 1) {
    uint16_t* bufs[2] = {k.d_work, k.d_work2 ? k.d_work2 : k.d_work};
    for (int b = 0; b < 2; ++b) {
      TRY(encode_tmap(&k.tmap16[b], CU_TENSOR_MAP_DATA_TYPE_UINT16, bufs[b], (uint64_t)k.W, (uint64_t)k.H, (uint64_t)c->batch,
                      (uint64_t)k.W * 2, (uint64_t)k.npix * 2, (uint32_t)g.box_px16, (uint32_t)g.box_rows, 1));
      TRY(encode_tmap(&k.tmap16g[b], CU_TENSOR_MAP_DATA_TYPE_UINT16, bufs[b], (uint64_t)k.W, (uint64_t)k.H, (uint64_t)c->batch,
                      (uint64_t)k.W * 2, (uint64_t)k.npix * 2, (uint32_t)g.box_px16, (uint32_t)g.box_rows, (uint32_t)g.group_frames));
    }
  } else {
    const uint64_t row_bytes = (uint64_t)k.W * 3 / 2;
    TRY(encode_tmap(&k.tmap12, CU_TENSOR_MAP_DATA_TYPE_UINT32, k.d_in, row_bytes / 4, (uint64_t)k.H, (uint64_t)c->capacity,
                    row_bytes, (uint64_t)k.frame_bytes, (uint32_t)g.box_words12, (uint32_t)g.box_rows, 1));
    TRY(encode_tmap(&k.tmap12g, CU_TENSOR_MAP_DATA_TYPE_UINT32, k.d_in, row_bytes / 4, (uint64_t)k.H, (uint64_t)c->capacity,
                    row_bytes, (uint64_t)k.frame_bytes, (uint32_t)g.box_words12, (uint32_t)g.box_rows, (uint32_t)g.group_frames));
    for (int b = 0; b < 2; ++b) {
      CU(cudaMalloc(&k.d_fix[b], (size_t)c->batch * hot_fix_bytes()));
      CU(cudaMemset(k.d_fix[b], 0, (size_t)c->batch * hot_fix_bytes()));
    }
  }
  // 16-bit node-major rows?  Unit projection values make every value of a plain node an integer < 2^16; the other nodes
  // (patched: float values, unseen: NaN) keep float rows in a side buffer behind the 16-bit rows of their owner.  Needs
  // the symmetric phase-2 kernel (the one that reads 16-bit rows) and 16-byte aligned row segments.  UPSP_ITRANS16=0: off.
  {
    const int cl = phase2_cluster(c->F, true);
    bool ok = val1 && !(getenv("UPSP_ITRANS16") && atoi(getenv("UPSP_ITRANS16")) == 0) && cl > 0 && c->F % (8 * cl) == 0 &&
              c->F % 8 == 0;
    // batch-blocked layout: several ranks, every rank the same number of frames (a multiple of 8), power-of-two batch
    int blk_len = 0, blk_kb = 0;
    {
      static const int blocked_env = getenv("UPSP_BLOCKED") ? atoi(getenv("UPSP_BLOCKED")) : -1;      // 0: row-major, 1: also on one rank
      const bool pow2 = c->batch >= 64 && (c->batch & (c->batch - 1)) == 0;
      bool equal = c->F % c->R == 0 && (c->F / c->R) % 8 == 0;
      for (int r = 0; r < c->R; ++r) equal = equal && c->f_count[r] == c->F / c->R && c->f_start[r] == r * (c->F / c->R);
      if (ok && pow2 && equal && blocked_env != 0 && (c->R > 1 || blocked_env == 1)) {
        blk_len = c->batch;
        blk_kb = (c->F / c->R + blk_len - 1) / blk_len;
      }
    }
    std::vector<int> oidx(N, -1), cnt(c->R, 0), local;
    for (int r = 0; r < c->R && ok; ++r) {
      for (int n = c->n_start[r]; n < c->n_start[r] + c->n_count[r]; ++n)
        if (k.h_code[n] < 0) {
          if (r == c->rank) local.push_back(n - c->n_start[r]);
          oidx[n] = cnt[r]++;
        }
      auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
      const size_t row_elems = blk_len > 0 ? (size_t)c->R * blk_kb * blk_len : (size_t)c->F;      // blocked: tail batches padded
      c->side_off[r] = al((size_t)c->n_count[r] * row_elems * sizeof(uint16_t));
      // the side buffer must fit behind the 16-bit rows inside the block sized for float rows
      ok = ok && c->side_off[r] + (size_t)cnt[r] * c->F * sizeof(float) <= (size_t)c->n_count[r] * c->F * sizeof(float);
    }
    if (ok) {
      c->it16 = true;
      c->blk_len = blk_len;
      c->blk_kb = blk_kb;
      c->n_other_local = (int)local.size();
      TRY(upload(&c->d_other_idx, oidx.data(), oidx.size()));
      if (!local.empty()) TRY(upload(&c->d_other_local, local.data(), local.size()));
      // rows go straight to their owners in this mode (half the NVLink bytes of the float rows): no staging block
      c->staged_xchg = false;
      c->staged_peers = 0;
      c->stage_mask = 0;
      for (int i = 0; i < 2; ++i) {
        cudaFree(c->d_stage[i]);
        c->d_stage[i] = nullptr;
      }
    }
  }
  c->proj_mode = want;
  return UPSP_OK;
}

// ------------------------------------------------------------------------------------------
// phase 1
// ------------------------------------------------------------------------------------------
static size_t frame_bytes_of(int format, size_t npix) {
  switch (format) {
    case UPSP_PIX_U16: return npix * 2;
    case UPSP_PIX_PACKED12: return npix * 12 / 8;
    case UPSP_PIX_PACKED10: return npix * 10 / 8;
  }
  return 0;
}

extern "C" int upsp_gpu_push_frames(upsp_gpu_ctx* c, int cam, const void* host, int format,
                                    int off, int count) {
  ENTER(c);
  CAM_CHECK(c, cam);
  Camera& k = c->cams[cam];
  REQUIRE(k.npix > 0, UPSP_ERR_STATE, "set_camera(%d) first", cam);
  REQUIRE(format >= UPSP_PIX_U16 && format <= UPSP_PIX_PACKED10, UPSP_ERR_INVALID, "format %d", format);
  REQUIRE(off >= 0 && count >= 0 && off + count <= c->F_local, UPSP_ERR_INVALID,
          "frames [%d,%d) outside the local slice of %d", off, off + count, c->F_local);
  REQUIRE(count <= c->capacity, UPSP_ERR_INVALID, "%d frames exceed the %d input slots", count, c->capacity);
  REQUIRE(host || count == 0, UPSP_ERR_INVALID, "null frames");
  if (format == UPSP_PIX_PACKED12)
    REQUIRE(k.npix % 2 == 0, UPSP_ERR_INVALID, "12-bit packing needs an even pixel count");
  if (format == UPSP_PIX_PACKED10)
    REQUIRE(k.npix % 4 == 0, UPSP_ERR_INVALID, "10-bit packing needs a pixel count divisible by 4");
  if (k.format < 0) {
    k.format = format;
    k.frame_bytes = frame_bytes_of(format, k.npix);
    TRY(dmalloc(&k.d_in, (size_t)c->capacity * k.frame_bytes));
  }
  REQUIRE(k.format == format, UPSP_ERR_INVALID, "camera %d was fed format %d before", cam, k.format);
  // slots being overwritten must have been consumed: wait for the process_frames calls
  // whose frames (an earlier lap of the ring) live in the slots of [off, off+count)
  {
    const long lo = (long)off - c->capacity, hi = (long)off + count - c->capacity;  // previous lap
    long covered = 0;           // frames of the previous lap whose process_frames record is still held
    for (auto& r : c->proc_recs) {
      if (r.count == 0) continue;
      const bool overlap_prev = r.off < hi && r.off + r.count > lo;
      const bool older = r.off + r.count <= lo;  // even older laps: also done by stream order, cheap to wait
      if (overlap_prev || older) CU(cudaStreamWaitEvent(c->copy_stream, r.ev, 0));
      if (overlap_prev) covered += std::min<long>(hi, r.off + r.count) - std::max<long>(lo, r.off);
    }
    // The record ring holds the last 16 calls.  When a lap of the input ring spans more calls than that, the
    // record of the slots being overwritten is gone: fall back to the most recent process_frames event (stream
    // order covers every older call).  Same after upsp_gpu_reset_run, which forgets the records.
    const long need = std::min<long>(hi, c->frames_processed) - std::max<long>(lo, 0);
    if (c->push_wait_all || (hi > 0 && covered < need)) CU(cudaStreamWaitEvent(c->copy_stream, c->ev_proc, 0));
    c->push_wait_all = false;
  }
  int done = 0;
  while (done < count) {
    const int slot = (off + done) % c->capacity;
    const int run = std::min(count - done, c->capacity - slot);
    CU(cudaMemcpyAsync(k.d_in + (size_t)slot * k.frame_bytes,
                       (const uint8_t*)host + (size_t)done * k.frame_bytes,
                       (size_t)run * k.frame_bytes, cudaMemcpyHostToDevice, c->copy_stream));
    done += run;
  }
  CU(cudaEventRecord(c->ev_push, c->copy_stream));
  return UPSP_OK;
}

template <int U>
static void launch_project_ell1(upsp_gpu_ctx* c, const ProjArgs& a) {
  const unsigned g = cdiv(a.n_nodes, 256);
  switch (a.n_cams) {
    case 1: k_project_ell1<1, U><<<g, 256, 0, c->stream>>>(a); break;
    case 2: k_project_ell1<2, U><<<g, 256, 0, c->stream>>>(a); break;
    case 3: k_project_ell1<3, U><<<g, 256, 0, c->stream>>>(a); break;
    case 4: k_project_ell1<4, U><<<g, 256, 0, c->stream>>>(a); break;
    case 5: k_project_ell1<5, U><<<g, 256, 0, c->stream>>>(a); break;
    case 6: k_project_ell1<6, U><<<g, 256, 0, c->stream>>>(a); break;
    case 7: k_project_ell1<7, U><<<g, 256, 0, c->stream>>>(a); break;
    default: k_project_ell1<8, U><<<g, 256, 0, c->stream>>>(a); break;
  }
}

// cv::findTransformECC for local frames [off, off+nb) of one camera (decoded frames in d_work);
// leaves the 2x3 maps in d_m6 and the per-frame outcome in d_eccState.
static int ecc_run_batch(upsp_gpu_ctx* c, Camera& k, int off, int nb) {
  const int max_iters = 50;       // psp_process.cpp:1779-1780
  const float eps = 0.001f;
  for (int s0 = 0; s0 < nb; s0 += ECC_BATCH) {
    const int n = std::min(ECC_BATCH, nb - s0);
    const uint16_t* fr = k.d_work + (size_t)s0 * k.npix;
    float* m6 = k.d_m6 + (size_t)(off + s0) * 6;
    EccState* st = k.d_eccState + off + s0;
    const dim3 g(cdiv(k.W, 256), k.H, n);
    k_ecc_blur_rows<uint16_t><<<g, 256, 0, c->stream>>>(fr, k.d_eccTmp, k.W, k.H);
    KCHECK(c);
    k_ecc_blur_cols<<<g, 256, 0, c->stream>>>(k.d_eccTmp, k.d_eccI, k.W, k.H);
    KCHECK(c);
    k_ecc_grad<<<g, 256, 0, c->stream>>>(k.d_eccI, k.d_eccG, k.W, k.H);
    KCHECK(c);
    // global frame 0 is never registered (psp_process.cpp:1777)
    const int skip = (c->f0 + off + s0 == 0) ? 0 : -1;
    k_ecc_init<<<cdiv(n, 64), 64, 0, c->stream>>>(st, m6, n, skip, eps);
    KCHECK(c);
    const int rows_per_block = (k.H + ECC_NBLK - 1) / ECC_NBLK;
    const int nblk = (k.H + rows_per_block - 1) / rows_per_block;
    for (int it = 1; it <= max_iters; ++it) {
      k_warp_tables<<<dim3(cdiv(std::max(k.W, k.H), 256), n), 256, 0, c->stream>>>(m6, n, k.W, k.H, 1, k.d_eccTab);
      KCHECK(c);
      CU(cudaMemsetAsync(k.d_nactive, 0, sizeof(int), c->stream));
      k_ecc_reduce<<<dim3(nblk, n), ECC_NT, 0, c->stream>>>(k.d_eccI, k.d_eccG, k.d_eccT, k.d_eccTab, st, k.W, k.H,
                                                            rows_per_block, k.d_eccPart);
      KCHECK(c);
      k_ecc_solve<<<cdiv(n, 32), 32, 0, c->stream>>>(k.d_eccPart, nblk, n, m6, st, max_iters, eps, k.d_nactive);
      KCHECK(c);
      if (it >= 2) {   // frames typically converge in 2-6 iterations: poll the survivor count
        int active = 0;
        CU(cudaMemcpyAsync(&active, k.d_nactive, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (active == 0) break;
      }
    }
  }
  return UPSP_OK;
}

// rows stored straight into peer memory (>= 3 ranks, or staging off, or 16-bit rows at any rank count): 128-byte segments.
// UPSP_FORCE_SEG128=1: the same kernel variant on one GPU (A/B measurements).
static bool seg128_variant(const upsp_gpu_ctx* c) {
  static const bool force = getenv("UPSP_FORCE_SEG128") && atoi(getenv("UPSP_FORCE_SEG128"));
  static const bool force64 = getenv("UPSP_FORCE_SEG64") && atoi(getenv("UPSP_FORCE_SEG64"));      // A/B knob: 64-byte segments to peers
  return ((c->R > 1 && c->staged_peers < c->R - 1) || force) && !force64;
}

// one batch: local frames [off, off+nb)
static int process_batch_impl(upsp_gpu_ctx* c, int off, int nb) {
  ProjArgs pa{};
  pa.n_cams = (int)c->cams.size();
  pa.n_nodes = c->N;
  pa.nframes = nb;
  pa.bstride = c->batch;
  pa.out = c->fused ? nullptr : c->d_intensity + (size_t)off * c->N;
  FusedArgs fa{};
  pa.sum = c->d_sum;
  pa.sumsq = c->d_sumsq;
  const int slot = off % c->capacity;
  REQUIRE(slot + nb <= c->capacity, UPSP_ERR_STATE, "batch wraps the input ring");
  for (size_t ci = 0; ci < c->cams.size(); ++ci)
    REQUIRE(c->cams[ci].format >= 0, UPSP_ERR_STATE, "camera %zu has no frames pushed", ci);
  TRY(ensure_proj_mode(c));
  const bool serial = c->sample_every > 0 && !c->timeline && (c->batch_counter++ % c->sample_every) == 0;
  const bool prof = serial || c->timeline;
  // pipeline: buffer set `bs`, front end on stream SB; it may start once the fused kernel that last
  // read this buffer set (two batches ago) is done
  // A batch whose kernels are being timed (upsp_gpu_set_kernel_sampling) runs un-overlapped: its
  // front end goes on the main stream, and the next batch's front end waits for its projection, so the
  // CUDA-event durations are those of the kernels alone (they are what the roofline figures use).
  const int bs = c->pipelined ? (int)(c->pipe_batches & 1) : 0;
  cudaStream_t SB = (c->pipelined && !serial) ? c->stream_b : c->stream;
  if (c->pipelined) {
    CU(cudaStreamWaitEvent(SB, c->ev_back[bs], 0));
    // UPSP_FRONT=serial: the front end of batch i+1 starts after the projection of batch i (no decode under the
    // projection; the patch kernel still runs beside the TMA kernel of its own batch)
    // Measured with the 128-byte-segment variant of the projection (rows stored straight into peer memory; 5 blocks per
    // SM): on ONE GPU with the variant forced, serial + a full-size scan wins (43.5 against 48.6 ms, r2z/r2aa); on 8 GPUs,
    // where the kernel also waits on its NVLink stores, the scan beside it fills those gaps and overlapped wins (51.5
    // against 54.7 ms, r2t/r2ac).  So overlapped stays the default everywhere.
    static const char* front_env = getenv("UPSP_FRONT");
    const bool front_serial = front_env ? !strcmp(front_env, "serial") : false;
    c->front_serial_now = front_serial;
    if (c->last_sampled || front_serial) CU(cudaStreamWaitEvent(SB, c->ev_back[bs ^ 1], 0));
    c->last_sampled = serial;
  }
  for (size_t ci = 0; ci < c->cams.size(); ++ci) {
    Camera& k = c->cams[ci];
    REQUIRE(k.format >= 0, UPSP_ERR_STATE, "camera %zu has no frames pushed", ci);
    uint16_t* const w_work = bs ? k.d_work2 : k.d_work;
    int* const w_hot_cnt = bs ? k.d_hot_cnt2 : k.d_hot_cnt;
    int* const w_hot_pos = bs ? k.d_hot_pos2 : k.d_hot_pos;
    float* const w_pv = bs ? k.d_pv2 : k.d_pv;
    const int thresh = c->hot_fix ? UPSP_HOT_THRESH : 0x7fffffff;
    CU(cudaMemsetAsync(w_hot_cnt, 0, (size_t)2 * c->batch * sizeof(int), SB));
    int* done = c->hot_fix ? w_hot_cnt + c->batch : nullptr;
    const uint8_t* in = k.d_in + (size_t)slot * k.frame_bytes;
    const bool src12 = c->proj_mode == 2;       // the projection reads the packed frames itself: scan only
    KBEGIN_ON(0, SB);
    if (src12) {
      if (c->hot_fix) {
        // one small block per SM beside the projection (measured r2g/r2l); a full-size grid when nothing runs beside it
        static const int scan_env = getenv("UPSP_SCAN_BPSM") ? atoi(getenv("UPSP_SCAN_BPSM")) : 0;
        const int scan_bpsm = scan_env > 0 ? scan_env : ((c->front_serial_now || serial || !c->pipelined) ? 4 : 1);
        const int scan_threads = (c->front_serial_now || serial || !c->pipelined) ? 256 : 128;
        CU(launch_hot_scan12(in, k.frame_bytes, k.npix, nb, thresh, w_hot_cnt, w_hot_pos, w_hot_cnt + c->batch, k.H, k.W,
                             k.d_fix[bs], c->n_sm * scan_bpsm, scan_threads, SB));
        KCHECK(c);
      }
    } else {
    static const int decode_p = getenv("UPSP_DECODE_P") ? atoi(getenv("UPSP_DECODE_P")) : -1;
    const bool persistent = (decode_p < 0 ? c->pipelined : decode_p != 0) && k.format == UPSP_PIX_PACKED12 &&
                            k.npix % 32 == 0 && k.frame_bytes % 16 == 0 && (k.npix / 32) * (size_t)nb < ((size_t)1 << 31);
    if (persistent) {
      static const int decode_bpsm = getenv("UPSP_DECODE_BPSM") ? atoi(getenv("UPSP_DECODE_BPSM")) : 2;
      k_unpack12_scan_p<<<c->n_sm * decode_bpsm, 256, 0, SB>>>(in, k.frame_bytes, w_work, k.npix, nb, thresh,
                                                              w_hot_cnt, w_hot_pos, done, k.H, k.W);
    } else if (k.format == UPSP_PIX_PACKED12) {
      k_unpack12_scan<<<dim3(cdiv(cdiv(k.npix, 8), 256), nb), 256, 0, SB>>>(
          in, k.frame_bytes, w_work, k.npix, thresh, w_hot_cnt, w_hot_pos, done, k.H, k.W);
    } else if (k.format == UPSP_PIX_PACKED10) {
      k_unpack10_scan<<<dim3(cdiv(cdiv(k.npix, 4), 256), nb), 256, 0, SB>>>(
          in, k.frame_bytes, w_work, k.npix, c->d_lut, thresh, w_hot_cnt, w_hot_pos, done, k.H, k.W);
    } else {
      k_copy16_scan<<<dim3(cdiv(cdiv(k.npix, 8), 256), nb), 256, 0, SB>>>(
          (const uint16_t*)in, k.npix, w_work, k.npix, thresh, w_hot_cnt, w_hot_pos, done, k.H, k.W);
    }
    KCHECK(c);
    }
    KEND_ON(SB);
    const bool reg = c->registration != UPSP_REG_NONE;
    if (c->registration == UPSP_REG_PIXEL) {
      TRY(ecc_run_batch(c, k, off, nb));
      k_warp_tables<<<dim3(cdiv(std::max(k.W, k.H), 256), nb), 256, 0, c->stream>>>(
          k.d_m6 + (size_t)off * 6, nb, k.W, k.H, c->interp, k.d_tab + (size_t)off * (2 * k.W + 2 * k.H));
      KCHECK(c);
      k_tma_coef<<<cdiv(nb, 256), 256, 0, c->stream>>>(k.d_m6 + (size_t)off * 6, nb, (c->f0 + off == 0) ? 0 : -1, k.d_coef + off);
      KCHECK(c);
    }
    if (c->proj_mode > 0)       // this batch's column coefficients -> constant table `bs`
      CU(tma_set_coef(bs, k.d_coef + off, nb, c->registration == UPSP_REG_PIXEL ? c->stream : SB));
    if (c->pipelined && c->proj_mode > 0) {   // the TMA kernel needs the decoded frames / fix lists, not the patch values
      CU(cudaEventRecord(c->ev_dec[bs], SB));
    }
    const int* tabs = reg ? k.d_tab + (size_t)off * (2 * k.W + 2 * k.H) : nullptr;   // this batch's tables
    const uint16_t* cur = w_work;
    // global frame 0 is never registered (psp_process.cpp:1777)
    const int skip_frame = (c->f0 + off == 0) ? 0 : -1;
    if (reg && !c->fused) {
      KBEGIN(2);
      k_warp_affine8_u16<<<dim3(cdiv(k.W, 1024), k.H, nb), 128, 0, c->stream>>>(
          w_work, k.d_warp, k.W, k.H, tabs, c->interp, skip_frame);
      KCHECK(c);
      KEND();
      cur = k.d_warp;
    }
    const bool patch = c->patcher == UPSP_PATCH_POLYNOMIAL && k.has_patches;
    if (patch) {
      const size_t sm = ((size_t)20 * k.max_bounds + PATCH_WARPS * ((size_t)2 * k.max_bounds + 16)) * sizeof(float);
      REQUIRE(sm <= 200 * 1024, UPSP_ERR_INVALID, "patch cluster with %d boundary pixels is too large", k.max_bounds);
      if (sm > 48 * 1024)
        CU(cudaFuncSetAttribute(k_patch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      KBEGIN_ON(3, SB);
      for (size_t l = 0; l + 1 < k.level_off.size(); ++l) {
        const int ncl = k.level_off[l + 1] - k.level_off[l];
        if (!ncl) continue;
        k_patch<<<dim3(ncl, cdiv(nb, PATCH_WARPS)), 32 * PATCH_WARPS, sm, SB>>>(
            k.geom, k.d_cl_list + k.level_off[l], cur, k.npix, k.W, k.H,
            (reg && c->fused) ? tabs : nullptr, c->interp, skip_frame, nb, c->batch, w_pv, src12 ? in : nullptr,
            k.frame_bytes, (src12 && c->hot_fix) ? reinterpret_cast<const HotFix*>(k.d_fix[bs]) : nullptr);
        KCHECK(c);
      }
      KEND_ON(SB);
    }
    const float* cur32 = nullptr;
    if (c->filter.kind) {   // only reachable in the unfused mode
      const dim3 fg(cdiv(k.W, 256), k.H, nb);
      if (patch) {
        const size_t tot = (size_t)nb * k.npix;
        k_u16_to_f32<<<cdiv(tot, 256), 256, 0, c->stream>>>(cur, k.d_img32, tot);
        KCHECK(c);
        k_scatter_patched<<<dim3(cdiv(k.total_internal, 256), nb), 256, 0, c->stream>>>(
            w_pv, k.d_slot_pix, k.total_internal, c->batch, nb, k.npix, k.d_img32);
        KCHECK(c);
        if (c->filter.kind == 1) {
          k_filter_f32<<<fg, 256, 0, c->stream>>>(k.d_img32, k.d_img32b, k.W, k.H, c->filter, 0);
          KCHECK(c);
          k_filter_f32<<<fg, 256, 0, c->stream>>>(k.d_img32b, k.d_img32, k.W, k.H, c->filter, 1);
          KCHECK(c);
          cur32 = k.d_img32;
        } else {
          k_filter_f32<<<fg, 256, 0, c->stream>>>(k.d_img32, k.d_img32b, k.W, k.H, c->filter, 0);
          KCHECK(c);
          cur32 = k.d_img32b;
        }
      } else {
        k_filter_u16<<<fg, 256, 0, c->stream>>>(cur, k.d_filt16, k.W, k.H, c->filter);
        KCHECK(c);
        cur = k.d_filt16;
      }
    }
    pa.cam[ci].frames = cur;
    pa.cam[ci].frames32 = cur32;
    pa.cam[ci].npix = k.npix;
    pa.cam[ci].pv = (patch && !c->filter.kind) ? w_pv : nullptr;
    pa.cam[ci].code = k.d_code;
    pa.cam[ci].val = k.d_val;
    pa.cam[ci].rowptr = k.d_rowptr;
    fa.cam[ci].frames = w_work;
    fa.cam[ci].npix = k.npix;
    fa.cam[ci].W = k.W;
    fa.cam[ci].H = k.H;
    fa.cam[ci].tab = tabs;
    fa.cam[ci].m6 = reg ? k.d_m6 + (size_t)off * 6 : nullptr;
    fa.cam[ci].pv = patch ? w_pv : nullptr;
    fa.cam[ci].code = k.d_code;
    fa.cam[ci].val = k.d_val;
    fa.skip_frame = skip_frame;
  }
  if (c->pipelined) {
    CU(cudaEventRecord(c->ev_front[bs], SB));
    // TMA modes: the big kernel starts on the decoded frames alone, the patch kernel runs beside it on the front-end
    // stream, and only the small kernel of the patched / unseen nodes waits for the patch values
    CU(cudaStreamWaitEvent(c->stream, c->proj_mode > 0 ? c->ev_dec[bs] : c->ev_front[bs], 0));
  }
  if (c->fused) {
    fa.n_cams = pa.n_cams;
    fa.n_nodes = c->N;
    fa.nframes = nb;
    fa.bstride = c->batch;
    fa.interp = c->interp;
    fa.sum = c->d_sum;
    fa.sumsq = c->d_sumsq;
    fa.perm = c->d_perm;
    fa.n_ranks = c->R;
    fa.f_total = c->F;
    fa.col0 = c->f0 + off;
    fa.rank = c->rank;
    fa.stage = c->staged_xchg ? c->d_stage[bs] : nullptr;
    fa.stage_stride = c->stage_stride;
    fa.stage_mask = c->stage_mask;
    if (c->staged_xchg)   // copies of two batches ago are out
      for (auto e : c->ev_x[bs]) CU(cudaStreamWaitEvent(c->stream, e, 0));
    for (int r = 0; r < c->R; ++r) {
      fa.dst[r] = reinterpret_cast<float*>(c->peer_base[r]);
      fa.node_start[r] = c->n_start[r];
    }
    fa.node_start[c->R] = c->N;
    const unsigned g = cdiv(c->N, 256);
    KBEGIN(4);
    const bool regk = c->registration != UPSP_REG_NONE;
    static const int fused_bs = getenv("UPSP_FUSED_BS") ? atoi(getenv("UPSP_FUSED_BS")) : 128;   // tuning knob: nodes per block
    bool int12 = true;   // every camera's container guarantees pixels < 2^14
    for (auto& k : c->cams)
      int12 = int12 && (k.format == UPSP_PIX_PACKED12 || (k.format == UPSP_PIX_PACKED10 && c->lut_max < 16384));
#define FUSED_LAUNCH(NCAM)                                                                          \
  if (regk && int12 && fused_bs == 64) k_project_fused<NCAM, true, true, 4, 64><<<cdiv(c->N, 64), 64, 0, c->stream>>>(fa); \
  else if (regk && int12 && fused_bs == 128) k_project_fused<NCAM, true, true, 4, 128><<<cdiv(c->N, 128), 128, 0, c->stream>>>(fa); \
  else if (regk && int12) k_project_fused<NCAM, true, true, 4, 256><<<g, 256, 0, c->stream>>>(fa);   \
  else if (regk) k_project_fused<NCAM, true, false, 4, 256><<<g, 256, 0, c->stream>>>(fa);           \
  else k_project_fused<NCAM, false, false, 8, 256><<<g, 256, 0, c->stream>>>(fa)
    // hot configuration (bilinear registration, 12-bit containers): k_project_fused4.  UPSP_FUSED_V1=1
    // keeps the first-generation kernel (A/B measurements, and the parity test that pins the two to
    // identical bits)
    static const bool fused_v1 = getenv("UPSP_FUSED_V1") && atoi(getenv("UPSP_FUSED_V1"));
    bool pix13 = true;   // pixels < 2^13: the bilinear sum fits under the float bit pattern of 2^23
    size_t max_elems = 0;
    for (auto& k : c->cams) {
      pix13 = pix13 && (k.format == UPSP_PIX_PACKED12 || (k.format == UPSP_PIX_PACKED10 && c->lut_max < 8192));
      max_elems = std::max(max_elems, (size_t)c->batch * k.npix);
      max_elems = std::max(max_elems, (size_t)c->batch * (size_t)(k.W + k.H));
    }
    if (c->proj_mode > 0) {
      // TMA-staged boxes for the nodes with a plain pixel, k_project_fused4 for the rest (patched / unseen nodes)
      Camera& k = c->cams[0];
      const bool seg128 = seg128_variant(c);
      fa.perm = c->d_perm_tma;
      TmaExtra ex{};
      ex.blk = c->d_tma_blk;
      ex.coef_set = bs;
      {
        // unit projection values (integer statistics): cut the batch into frame slices so that the grid's last wave is
        // short; UPSP_TMA_SPLIT = number of slices (default 2, 1 = off)
        static const int nsplit = getenv("UPSP_TMA_SPLIT") ? std::max(1, atoi(getenv("UPSP_TMA_SPLIT"))) : 2;
        const int stage = tma_stage_frames();
        const int per = ((nb + nsplit - 1) / nsplit + stage - 1) / stage * stage;
        ex.split_frames = (c->tma_val1 && nsplit > 1 && per < nb) ? per : 0;
      }
      if (c->proj_mode == 2) {
        ex.packed = k.d_in + (size_t)slot * k.frame_bytes;
        ex.frame_bytes = k.frame_bytes;
        ex.frame0 = slot;
        ex.hot = c->hot_fix ? reinterpret_cast<const HotFix*>(k.d_fix[bs]) : nullptr;
        fa.cam[0].frames = nullptr;
      }
      if (c->it16 && c->blk_len > 0) {
        fa.blk_len = c->blk_len;
        fa.blk_index = c->rank * c->blk_kb + off / c->blk_len;
        fa.blk_j0 = off % c->blk_len;
      }
      CU(launch_project_tma(c->proj_mode - 1, seg128, c->tma_val1, c->it16, c->proj_mode == 2 ? k.tmap12g : k.tmap16g[bs],
                            c->proj_mode == 2 ? k.tmap12 : k.tmap16[bs], fa, ex, c->n_tma_blocks, c->stream));
      c->launches++;
      const int n_other = c->N - c->n_tma_plain;
      if (c->pipelined) CU(cudaStreamWaitEvent(c->stream, c->ev_front[bs], 0));
      if (n_other > 0) {
        FusedArgs fb = fa;
        fb.perm = c->d_perm_tma + c->n_tma_plain;
        fb.n_nodes = n_other;
        if (c->it16) {      // float rows of the patched / unseen nodes: the owners' side buffers (node-major)
          fb.blk_len = 0;
          fb.row_index = c->d_other_idx;
          for (int r = 0; r < c->R; ++r) fb.dst[r] = reinterpret_cast<float*>(c->peer_base[r] + c->side_off[r]);
        }
        if (seg128) k_project_fused4<1, 128, 32><<<cdiv(n_other, 128), 128, 0, c->stream>>>(fb);
        else k_project_fused4<1, 128, 16><<<cdiv(n_other, 128), 128, 0, c->stream>>>(fb);
        KCHECK(c);
      }
      KEND();
      return UPSP_OK;
    }
    if (!fused_v1 && regk && pix13 && c->interp == UPSP_INTERP_LINEAR && max_elems < ((size_t)1 << 31)) {
      const unsigned g2 = cdiv(c->N, 128);
      // rows stored straight into peer memory (>= 3 ranks, or staging off): 128-byte segments
      const bool seg128 = c->R > 1 && c->staged_peers < c->R - 1;
      // UPSP_PROJ_PAD: extra dynamic shared memory per block = a cap on the resident projection blocks per SM,
      // which leaves registers for the co-resident front end (tuning knob)
      static const int proj_pad = getenv("UPSP_PROJ_PAD") ? atoi(getenv("UPSP_PROJ_PAD")) : 0;
#define FUSED4(NCAM)                                                               \
  if (seg128) k_project_fused4<NCAM, 128, 32><<<g2, 128, proj_pad, c->stream>>>(fa);      \
  else k_project_fused4<NCAM, 128, 16><<<g2, 128, proj_pad, c->stream>>>(fa)
      switch (fa.n_cams) {
        case 1: FUSED4(1); break;
        case 2: FUSED4(2); break;
        case 3: FUSED4(3); break;
        case 4: FUSED4(4); break;
        case 5: FUSED4(5); break;
        case 6: FUSED4(6); break;
        case 7: FUSED4(7); break;
        default: FUSED4(8); break;
      }
#undef FUSED4
      KCHECK(c);
      KEND();
      return UPSP_OK;
    }
    switch (fa.n_cams) {
      case 1: FUSED_LAUNCH(1); break;
      case 2: FUSED_LAUNCH(2); break;
      case 3: FUSED_LAUNCH(3); break;
      case 4: FUSED_LAUNCH(4); break;
      case 5: FUSED_LAUNCH(5); break;
      case 6: FUSED_LAUNCH(6); break;
      case 7: FUSED_LAUNCH(7); break;
      default: FUSED_LAUNCH(8); break;
    }
#undef FUSED_LAUNCH
    KCHECK(c);
    KEND();
    return UPSP_OK;
  }
  KBEGIN(4);
  if (c->ell1)
    launch_project_ell1<4>(c, pa);
  else
    k_project_csr<4><<<cdiv(c->N, 256), 256, 0, c->stream>>>(pa);
  KCHECK(c);
  KEND();
  return UPSP_OK;
}

static int process_batch(upsp_gpu_ctx* c, int off, int nb) {
  TRY(process_batch_impl(c, off, nb));
  if (c->pipelined) {   // the projection of this batch is the last reader of its buffer set
    const int bs = (int)(c->pipe_batches & 1);
    CU(cudaEventRecord(c->ev_back[bs], c->stream));
    if (c->staged_xchg) {
      // ship the other ranks' rows: [N_s rows x nb frames] block of the staging buffer -> columns
      // [f0+off, f0+off+nb) of rank s's node-major buffer, one strided copy per peer (copy engines)
      constexpr int NX = upsp_gpu_ctx::NX;
      for (int j = 0; j < NX; ++j) CU(cudaStreamWaitEvent(c->stream_x[j], c->ev_back[bs], 0));
      if (c->ship_sm) {
        // SM shipper (kernels_transpose.cuh): one small kernel for all staged peers, beside the next projection
        ShipArgs sa{};
        sa.stage = c->d_stage[bs];
        sa.stride = c->stage_stride;
        sa.nb = nb;
        sa.dst_stride = (size_t)c->F;
        for (int d = 1; d <= c->staged_peers; ++d) {
          const int s = (c->rank + d) % c->R;
          if (c->n_count[s] <= 0) continue;
          sa.dst[sa.n_peers] = reinterpret_cast<float*>(c->peer_base[s]) + (size_t)(c->f0 + off);
          sa.row0[sa.n_peers] = c->n_start[s];
          sa.rows[sa.n_peers] = c->n_count[s];
          sa.max_rows = std::max(sa.max_rows, c->n_count[s]);
          sa.n_peers++;
        }
        if (sa.n_peers > 0) {
          k_ship_rows<<<c->n_sm * c->ship_bpsm, 128, 0, c->stream_x[0]>>>(sa);
          KCHECK(c);
        }
      } else {
      // every peer's row block is cut into `parts` pieces so that NX copies are in flight at any time
      const int parts = std::max(1, (NX + c->staged_peers - 1) / c->staged_peers);
      int q = 0;
      for (int d = 1; d <= c->staged_peers; ++d) {
        const int s = (c->rank + d) % c->R;
        for (int part = 0; part < parts; ++part) {
          const int r0 = (int)((long long)c->n_count[s] * part / parts), r1 = (int)((long long)c->n_count[s] * (part + 1) / parts);
          if (r1 <= r0) continue;
          float* dst = reinterpret_cast<float*>(c->peer_base[s]) + (size_t)r0 * c->F + (size_t)(c->f0 + off);
          const float* src = c->d_stage[bs] + (size_t)(c->n_start[s] + r0) * c->stage_stride;
          CU(cudaMemcpy2DAsync(dst, (size_t)c->F * sizeof(float), src, (size_t)c->stage_stride * sizeof(float),
                               (size_t)nb * sizeof(float), (size_t)(r1 - r0), cudaMemcpyDeviceToDevice,
                               c->stream_x[q++ % NX]));
        }
      }
      }
      for (int j = 0; j < NX; ++j) CU(cudaEventRecord(c->ev_x[bs][j], c->stream_x[j]));
    }
    c->pipe_batches++;
  }
  return UPSP_OK;
}

extern "C" int upsp_gpu_process_frames(upsp_gpu_ctx* c, int off, int count) {
  ENTER(c);
  REQUIRE(off >= 0 && count >= 0 && off + count <= c->F_local, UPSP_ERR_INVALID,
          "frames [%d,%d) outside the local slice of %d", off, off + count, c->F_local);
  TRY(finalize(c));
  if (c->fused && c->R > 1)
    REQUIRE(c->peers_ready, UPSP_ERR_STATE,
            "multi-rank context is not wired (upsp_gpu_ipc_import / upsp_gpu_connect_local): the fused "
            "projection stores node-major rows straight into peer buffers");
  CU(cudaStreamWaitEvent(c->stream, c->ev_push, 0));
  CU(cudaEventRecord(c->ev_pa, c->stream));
  if (c->registration == UPSP_REG_GIVEN && count > 0) {
    // OpenCV-style fixed-point warp tables of every frame of this call (one launch per camera)
    for (auto& k : c->cams) {
      for (int s0 = 0; s0 < count; s0 += 65535) {      // grid.y limit: slabs of frames
        const int ns = std::min(65535, count - s0);
        k_warp_tables<<<dim3(cdiv(std::max(k.W, k.H), 256), ns), 256, 0, c->stream>>>(
            k.d_m6 + (size_t)(off + s0) * 6, ns, k.W, k.H, c->interp, k.d_tab + (size_t)(off + s0) * (2 * k.W + 2 * k.H));
        KCHECK(c);
      }
      const int skip = (c->f0 + off <= 0 && c->f0 + off + count > 0) ? -(c->f0 + off) : -1;
      k_tma_coef<<<cdiv(count, 256), 256, 0, c->stream>>>(k.d_m6 + (size_t)off * 6, count, skip, k.d_coef + off);
      KCHECK(c);
    }
  }
  if (c->pipelined) {   // the front-end stream needs the pushed frames and (patch kernel) the warp tables
    CU(cudaStreamWaitEvent(c->stream_b, c->ev_push, 0));
    CU(cudaEventRecord(c->ev_tabs, c->stream));
    CU(cudaStreamWaitEvent(c->stream_b, c->ev_tabs, 0));
  }
  int done = 0;
  while (done < count) {
    const int o = off + done;
    int nb = std::min(c->batch, count - done);
    nb = std::min(nb, c->capacity - (o % c->capacity));
    if (c->proj_mode < 0 && done == 0) nb = std::min(nb, c->batch - (o % c->batch));      // mode not decided yet: stay inside a block
    if (c->blk_len > 0) nb = std::min(nb, c->blk_len - (o % c->blk_len));                 // blocked rows: a batch lives in one block
    TRY(process_batch(c, o, nb));
    done += nb;
  }
  if (c->staged_xchg) {   // the call is complete when the last two batches' rows have left
    for (int i = 0; i < 2; ++i)
      for (auto e : c->ev_x[i]) CU(cudaStreamWaitEvent(c->stream, e, 0));
  }
  CU(cudaEventRecord(c->ev_pb, c->stream));
  CU(cudaEventRecord(c->ev_proc, c->stream));
  {
    auto& r = c->proc_recs[c->proc_next++ % c->proc_recs.size()];
    r.off = off;
    r.count = count;
    CU(cudaEventRecord(r.ev, c->stream));
  }
  c->frames_processed += count;
  c->stage_ms[0] = -1.0f;  // resolved lazily in stage_ms (needs a sync)
  return UPSP_OK;
}

// sum over ranks in rank order: identical bits on every rank
__global__ void k_allreduce_peer(double* const* bases_sum, double* const* bases_sq, int n_ranks,
                                 int n, double* out_sum, double* out_sq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0, q = 0.0;
  for (int r = 0; r < n_ranks; ++r) {
    s += bases_sum[r][i];
    q += bases_sq[r][i];
  }
  out_sum[i] = s;
  out_sq[i] = q;
}

extern "C" int upsp_gpu_finish_phase1(upsp_gpu_ctx* c) {
  ENTER(c);
  REQUIRE(c->finalized, UPSP_ERR_STATE, "no frames processed");
  if (c->registration == UPSP_REG_PIXEL && c->F_local > 0) {
    // the reference dies with cv::Exception when ECC does not converge (registration.cpp:64)
    CU(cudaStreamSynchronize(c->stream));
    std::vector<EccState> st(c->F_local);
    for (size_t ci = 0; ci < c->cams.size(); ++ci) {
      CU(cudaMemcpy(st.data(), c->cams[ci].d_eccState, st.size() * sizeof(EccState), cudaMemcpyDeviceToHost));
      for (int f = 0; f < c->F_local; ++f)
        REQUIRE(st[f].status != 2, UPSP_ERR_NUMERIC,
                "ECC registration failed (NaN correlation or lambda_d <= 0) for camera %zu, frame %d", ci, c->f0 + f);
    }
  }
  CU(cudaEventRecord(c->ev_a, c->stream));
  const double *sum = c->d_sum, *sq = c->d_sumsq;
  if (c->R > 1 && c->exchange == UPSP_XCHG_NCCL) {
    // the reference's MPI_Allreduce(MPI_SUM) of the per-rank sums (psp_process.cpp:1866-1872)
    REQUIRE(c->nccl_comm != nullptr, UPSP_ERR_STATE, "UPSP_XCHG_NCCL: call upsp_gpu_nccl_init first");
    NCCLCHK(nccl().AllReduce(c->d_sum, c->d_tsum, (size_t)c->N, NcclApi::kDouble, NcclApi::kSum, c->nccl_comm, c->stream));
    NCCLCHK(nccl().AllReduce(c->d_sumsq, c->d_tsq, (size_t)c->N, NcclApi::kDouble, NcclApi::kSum, c->nccl_comm, c->stream));
    sum = c->d_tsum;
    sq = c->d_tsq;
  } else if (c->R > 1) {
    REQUIRE(c->peers_ready, UPSP_ERR_STATE,
            "multi-rank context is not wired (upsp_gpu_ipc_import / upsp_gpu_connect_local)");
    // every rank lays its shared block out as [itrans N_r x F | sum N | sumsq N]: offsets depend on N_r
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    std::vector<double*> ptrs(2 * c->R);
    for (int r = 0; r < c->R; ++r) {
      const size_t osum = al((size_t)c->n_count[r] * c->F * sizeof(float));
      const size_t osq = osum + al((size_t)c->N * sizeof(double));
      ptrs[r] = reinterpret_cast<double*>(c->peer_base[r] + osum);
      ptrs[c->R + r] = reinterpret_cast<double*>(c->peer_base[r] + osq);
    }
    CU(cudaMemcpyAsync(c->d_peer_ptrs, ptrs.data(), ptrs.size() * sizeof(double*), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));   // ptrs is a stack object
    CU(cudaEventRecord(c->ev_a, c->stream));
    k_allreduce_peer<<<cdiv(c->N, 256), 256, 0, c->stream>>>(c->d_peer_ptrs, c->d_peer_ptrs + c->R, c->R, c->N,
                                                             c->d_tsum, c->d_tsq);
    KCHECK(c);
    sum = c->d_tsum;
    sq = c->d_tsq;
  }
  k_phase1_finals<<<cdiv(c->N, 256), 256, 0, c->stream>>>(sum, sq, c->N, (unsigned)c->F, c->d_avg, c->d_rms);
  KCHECK(c);
  CU(cudaEventRecord(c->ev_b, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaEventElapsedTime(&c->stage_ms[1], c->ev_a, c->ev_b));
  c->phase1_done = true;
  return UPSP_OK;
}

extern "C" int upsp_gpu_transpose(upsp_gpu_ctx* c) {
  ENTER(c);
  REQUIRE(c->finalized, UPSP_ERR_STATE, "no frames processed");
  if (c->exchange == UPSP_XCHG_NCCL && c->R > 1) {
    // The reference's own structure (local_transpose + global_transpose, psp_process.cpp:647-771) on NCCL: per chunk of
    // this rank's frames, k_transpose_a2a packs one node-major block [N_s][chunk] per destination rank into the send
    // buffer, one grouped ncclSend / ncclRecv round moves the blocks, and strided device copies reassemble the received
    // blocks into columns [f0_r + c0, ...) of intensity_transpose.  Chunked so that the send / receive buffers stay small
    // (N x chunk floats each); every rank derives every rank's chunk sizes from apportion(), so the rounds match.
    REQUIRE(c->nccl_comm != nullptr, UPSP_ERR_STATE, "UPSP_XCHG_NCCL: call upsp_gpu_nccl_init first");
    REQUIRE(!c->fused, UPSP_ERR_STATE, "UPSP_XCHG_NCCL needs the frame-major intermediate");
    const int Fc = 1024;
    int fmax = 0;
    for (int r = 0; r < c->R; ++r) fmax = std::max(fmax, c->f_count[r]);
    if (!c->d_nccl_send) {
      TRY(dmalloc(&c->d_nccl_send, (size_t)c->N * Fc));
      TRY(dmalloc(&c->d_nccl_recv, (size_t)std::max(c->N_local, 1) * Fc * c->R));
    }
    CU(cudaEventRecord(c->ev_a, c->stream));
    for (int c0 = 0; c0 < fmax; c0 += Fc) {
      const int nc = std::max(0, std::min(Fc, c->F_local - c0));
      if (nc > 0) {
        XposeArgs a{};
        a.src = c->d_intensity + (size_t)c0 * c->N;
        a.rows = nc;
        a.cols = c->N;
        a.n_ranks = c->R;
        a.f_total = nc;
        a.col0 = 0;
        for (int r = 0; r < c->R; ++r) {
          a.dst[r] = c->d_nccl_send + (size_t)c->n_start[r] * nc;
          a.node_start[r] = c->n_start[r];
        }
        a.node_start[c->R] = c->N;
        k_transpose_a2a<<<dim3(cdiv(c->N, XT), cdiv(nc, XT)), 256, 0, c->stream>>>(a);
        KCHECK(c);
      }
      NCCLCHK(nccl().GroupStart());
      size_t roff = 0;
      std::vector<size_t> roffs(c->R);
      for (int r = 0; r < c->R; ++r) {
        const int ncr = std::max(0, std::min(Fc, c->f_count[r] - c0));
        roffs[r] = roff;
        if (nc > 0 && c->n_count[r] > 0)
          NCCLCHK(nccl().Send(c->d_nccl_send + (size_t)c->n_start[r] * nc, (size_t)c->n_count[r] * nc, NcclApi::kFloat, r,
                              c->nccl_comm, c->stream));
        if (ncr > 0 && c->N_local > 0)
          NCCLCHK(nccl().Recv(c->d_nccl_recv + roff, (size_t)c->N_local * ncr, NcclApi::kFloat, r, c->nccl_comm, c->stream));
        roff += (size_t)c->N_local * ncr;
      }
      NCCLCHK(nccl().GroupEnd());
      for (int r = 0; r < c->R; ++r) {      // reassembly (psp_process.cpp:755-765)
        const int ncr = std::max(0, std::min(Fc, c->f_count[r] - c0));
        if (ncr > 0 && c->N_local > 0)
          CU(cudaMemcpy2DAsync(c->d_itrans + (size_t)c->f_start[r] + c0, (size_t)c->F * sizeof(float), c->d_nccl_recv + roffs[r],
                               (size_t)ncr * sizeof(float), (size_t)ncr * sizeof(float), (size_t)c->N_local,
                               cudaMemcpyDeviceToDevice, c->stream));
      }
    }
    CU(cudaEventRecord(c->ev_b, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventElapsedTime(&c->stage_ms[2], c->ev_a, c->ev_b));
    c->transposed = true;
    return UPSP_OK;
  }
  if (c->fused) {  // k_project_fused already wrote node-major rows (local and peer)
    CU(cudaStreamSynchronize(c->stream));
    c->stage_ms[2] = 0.0f;
    c->transposed = true;
    return UPSP_OK;
  }
  if (c->R > 1)
    REQUIRE(c->peers_ready, UPSP_ERR_STATE,
            "multi-rank context is not wired (upsp_gpu_ipc_import / upsp_gpu_connect_local)");
  XposeArgs a{};
  a.src = c->d_intensity;
  a.rows = c->F_local;
  a.cols = c->N;
  a.n_ranks = c->R;
  a.f_total = c->F;
  a.col0 = c->f0;
  for (int r = 0; r < c->R; ++r) {
    a.dst[r] = reinterpret_cast<float*>(c->peer_base[r]);
    a.node_start[r] = c->n_start[r];
  }
  a.node_start[c->R] = c->N;
  CU(cudaEventRecord(c->ev_a, c->stream));
  const bool prof = c->sample_every > 0;
  if (c->F_local > 0) {
    KBEGIN(5);
    k_transpose_a2a<<<dim3(cdiv(c->N, XT), cdiv(c->F_local, XT)), 256, 0, c->stream>>>(a);
    KCHECK(c);
    KEND();
  }
  CU(cudaEventRecord(c->ev_b, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaEventElapsedTime(&c->stage_ms[2], c->ev_a, c->ev_b));
  c->transposed = true;
  return UPSP_OK;
}

// ------------------------------------------------------------------------------------------
// phase 2
// ------------------------------------------------------------------------------------------
// inverse Gram matrix of T_k(x_f), x_f as the device computes it, in long double
static void cheb_ginv(int F, int nc, float xa, float xb, bool sym, double* ginv) {
  std::vector<long double> G((size_t)nc * nc, 0.0L);
  std::vector<long double> T(nc);
  for (int f = 0; f < F; ++f) {
    // symmetric kernel: the abscissa of a mirrored sample is minus its partner's
    const float xf = (sym && f >= F / 2) ? -fmaf((float)(F - 1 - f), xa, xb) : fmaf((float)f, xa, xb);
    const long double x = (long double)xf;
    T[0] = 1.0L;
    if (nc > 1) T[1] = x;
    for (int k = 2; k < nc; ++k) T[k] = 2.0L * x * T[k - 1] - T[k - 2];
    for (int i = 0; i < nc; ++i)
      for (int j = 0; j < nc; ++j) G[(size_t)i * nc + j] += T[i] * T[j];
  }
  // Gauss-Jordan with partial pivoting
  std::vector<long double> I((size_t)nc * nc, 0.0L);
  for (int i = 0; i < nc; ++i) I[(size_t)i * nc + i] = 1.0L;
  for (int col = 0; col < nc; ++col) {
    int piv = col;
    for (int r = col + 1; r < nc; ++r)
      if (fabsl(G[(size_t)r * nc + col]) > fabsl(G[(size_t)piv * nc + col])) piv = r;
    if (G[(size_t)piv * nc + col] == 0.0L) continue;  // F < nc: rank deficient, leave zero rows
    if (piv != col)
      for (int j = 0; j < nc; ++j) {
        std::swap(G[(size_t)piv * nc + j], G[(size_t)col * nc + j]);
        std::swap(I[(size_t)piv * nc + j], I[(size_t)col * nc + j]);
      }
    const long double d = G[(size_t)col * nc + col];
    for (int j = 0; j < nc; ++j) {
      G[(size_t)col * nc + j] /= d;
      I[(size_t)col * nc + j] /= d;
    }
    for (int r = 0; r < nc; ++r) {
      if (r == col) continue;
      const long double m = G[(size_t)r * nc + col];
      if (m == 0.0L) continue;
      for (int j = 0; j < nc; ++j) {
        G[(size_t)r * nc + j] -= m * G[(size_t)col * nc + j];
        I[(size_t)r * nc + j] -= m * I[(size_t)col * nc + j];
      }
    }
  }
  // fold in the Chebyshev -> power basis change: T_k(x) = sum_j C[k][j] x^j, so the power-basis
  // coefficients of the fit are p = C^T (Ginv m)
  std::vector<long double> Cm((size_t)nc * nc, 0.0L);
  Cm[0] = 1.0L;
  if (nc > 1) Cm[(size_t)1 * nc + 1] = 1.0L;
  for (int k = 2; k < nc; ++k)
    for (int j = 0; j < nc; ++j)
      Cm[(size_t)k * nc + j] = (j > 0 ? 2.0L * Cm[(size_t)(k - 1) * nc + j - 1] : 0.0L) - Cm[(size_t)(k - 2) * nc + j];
  for (int j = 0; j < nc; ++j)
    for (int i = 0; i < nc; ++i) {
      long double t = 0.0L;
      for (int k = 0; k < nc; ++k) t += Cm[(size_t)k * nc + j] * I[(size_t)k * nc + i];
      ginv[(size_t)j * nc + i] = (double)t;
    }
}

static int phase2_cluster(int F, bool in16);
static bool phase2_symmetric(const Phase2Args& a);

template <int NC, bool ROW_SMEM, int CL, int NT>
static int launch_phase2_cl_nt(const Phase2Args& a, size_t smem, cudaStream_t st);

// Short rows (<= 8192 frames, one CTA per row): 128 threads per row instead of 512.  The per-row set-up (block
// reductions, coefficient solve) is a quarter of the warp instructions then, and several rows share an SM.
template <int NC, bool ROW_SMEM, int CL>
static int launch_phase2_cl(const Phase2Args& a, size_t smem, cudaStream_t st) {
  if constexpr (CL == 1 && ROW_SMEM) {
    if (a.F <= 8192) return launch_phase2_cl_nt<NC, ROW_SMEM, CL, 128>(a, smem, st);
  }
  return launch_phase2_cl_nt<NC, ROW_SMEM, CL, 512>(a, smem, st);
}

template <int NC, bool ROW_SMEM, int CL, int NT>
static int launch_phase2_cl_nt(const Phase2Args& a, size_t smem, cudaStream_t st) {
  auto kern = k_phase2<NC, ROW_SMEM, NT, CL>;
  // static + dynamic shared memory may exceed the 48 KB default even when the dynamic part alone does not
  if (smem > 24 * 1024)
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)a.n_local * CL);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  CU(cudaLaunchKernelEx(&cfg, kern, a));
  return UPSP_OK;
}

template <int NC, int CL, bool PK, bool IN16 = false>
static int launch_phase2_sym_pk(const Phase2Args& a, cudaStream_t st);

// UPSP_PHASE2_SCALAR=1 keeps the scalar kernel (A/B measurements; the two must agree to the last bit except where the
// packed one sums a thread's Chebyshev moments in two interleaved partial sums)
template <int NC, int CL>
static int launch_phase2_sym(const Phase2Args& a, cudaStream_t st) {
  static const bool scalar = getenv("UPSP_PHASE2_SCALAR") && atoi(getenv("UPSP_PHASE2_SCALAR"));
  if (a.itrans16 != nullptr) return launch_phase2_sym_pk<NC, CL, true, true>(a, st);
  return scalar ? launch_phase2_sym_pk<NC, CL, false>(a, st) : launch_phase2_sym_pk<NC, CL, true>(a, st);
}

template <int NC, int CL, bool PK, bool IN16, int NT>
static int launch_phase2_sym_nt(const Phase2Args& a, cudaStream_t st);

// threads per CTA: 512; 128 for short rows (see launch_phase2_cl); 256 for clustered 16-bit rows (see phase2_cluster).
// UPSP_P2_NT=512 keeps 512 threads for those (A/B).
template <int NC, int CL, bool PK, bool IN16>
static int launch_phase2_sym_pk(const Phase2Args& a, cudaStream_t st) {
  if constexpr (CL == 1 && PK) {
    if (a.F <= 8192) return launch_phase2_sym_nt<NC, CL, PK, IN16, 128>(a, st);
  }
  if constexpr (PK && IN16 && CL >= 2) {
    static const int nt_env = getenv("UPSP_P2_NT") ? atoi(getenv("UPSP_P2_NT")) : 0;
    if (nt_env != 512) return launch_phase2_sym_nt<NC, CL, PK, IN16, 256>(a, st);
  }
  return launch_phase2_sym_nt<NC, CL, PK, IN16, 512>(a, st);
}

template <int NC, int CL, bool PK, bool IN16, int NT>
static int launch_phase2_sym_nt(const Phase2Args& a, cudaStream_t st) {
  auto kern = k_phase2_sym<NC, NT, CL, PK, IN16>;
  const size_t smem = (size_t)(a.F / CL) * sizeof(float);
  // static + dynamic shared memory may exceed the 48 KB default even when the dynamic part alone does not
  if (smem > 24 * 1024)
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)a.n_local * CL);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  CU(cudaLaunchKernelEx(&cfg, kern, a));
  return UPSP_OK;
}

template <int NC>
static int launch_phase2(upsp_gpu_ctx* c, const Phase2Args& a, cudaStream_t st, long long* launches) {
  (void)c;
  if (a.n_local == 0) return UPSP_OK;
  const int cl = phase2_cluster(a.F, a.itrans16 != nullptr);
  int rc = UPSP_OK;
  REQUIRE(phase2_symmetric(a) || (a.itrans16 == nullptr && a.row_list == nullptr), UPSP_ERR_STATE,
          "16-bit intensity rows need the symmetric phase-2 kernel");
  if (phase2_symmetric(a)) {
    Phase2Args b = a;
    if (cl > 1) {      // clustered rows leave their partial statistics in global memory (one cluster barrier per row)
      const size_t need = ((size_t)c->N_local + 1) * 8 * 4;
      if (c->cl_parts_n < need) {
        cudaFree(c->d_cl_parts);
        c->d_cl_parts = nullptr;
        TRY(dmalloc(&c->d_cl_parts, need));
        c->cl_parts_n = need;
      }
      b.cl_parts = c->d_cl_parts;
    }
    switch (cl) {
      case 1: rc = launch_phase2_sym<NC, 1>(b, st); break;
      case 2: rc = launch_phase2_sym<NC, 2>(b, st); break;
      case 4: rc = launch_phase2_sym<NC, 4>(b, st); break;
      default: rc = launch_phase2_sym<NC, 8>(b, st); break;
    }
    if (!rc && cl > 1) {
      k_phase2_cl_parts<<<cdiv(b.n_local, 256), 256, 0, st>>>(b, cl);
      ++*launches;
    }
  } else if (cl > 0) {
    const size_t seg = (size_t)(((a.F + cl * 4 - 1) / (cl * 4)) * 4) * sizeof(float);
    switch (cl) {
      case 1: rc = launch_phase2_cl<NC, true, 1>(a, seg, st); break;
      case 2: rc = launch_phase2_cl<NC, true, 2>(a, seg, st); break;
      case 4: rc = launch_phase2_cl<NC, true, 4>(a, seg, st); break;
      default: rc = launch_phase2_cl<NC, true, 8>(a, seg, st); break;
    }
  } else {
    rc = launch_phase2_cl<NC, false, 1>(a, 0, st);   // two HBM passes
  }
  if (rc) return rc;
  ++*launches;
  CU(cudaGetLastError());
  return UPSP_OK;
}

// long 16-bit rows: the streaming kernel pair (kernels_phase2.cuh)
template <int NC>
static int launch_phase2_stream(upsp_gpu_ctx* c, const Phase2Args& a, cudaStream_t st, long long* launches) {
  if (a.n_local == 0) return UPSP_OK;
  const int nchunk = (a.F / 2 + P2S_CHUNK - 1) / P2S_CHUNK;
  const size_t ncoef = (size_t)a.n_local * (UPSP_MAX_COEF + 1), nparts = (size_t)a.n_local * nchunk * 2;
  if (c->p2coef_n < ncoef) {
    cudaFree(c->d_p2coef);
    c->d_p2coef = nullptr;
    TRY(dmalloc(&c->d_p2coef, ncoef));
    c->p2coef_n = ncoef;
  }
  if (c->p2parts_n < nparts) {
    cudaFree(c->d_p2parts);
    c->d_p2parts = nullptr;
    TRY(dmalloc(&c->d_p2parts, nparts));
    c->p2parts_n = nparts;
  }
  k_phase2_moments<NC, 512><<<a.n_local, 512, 0, st>>>(a, c->d_p2coef);
  CU(cudaGetLastError());
  k_phase2_apply<NC, 256><<<dim3(nchunk, a.n_local), 256, 0, st>>>(a, c->d_p2coef, c->d_p2parts, nchunk);
  CU(cudaGetLastError());
  k_phase2_parts<<<cdiv(a.n_local, 256), 256, 0, st>>>(a, c->d_p2parts, nchunk);
  CU(cudaGetLastError());
  *launches += 3;
  return UPSP_OK;
}

static int dispatch_phase2(upsp_gpu_ctx* c, const Phase2Args& a, cudaStream_t st, long long* launches) {
  // UPSP_PHASE2_STREAM = 1 forces the streaming pair (16-bit rows only; off by default, kept as the measured alternative
  // and for the parity test that pins it).  Measured (r2u, 1 GPU, 20 000-frame rows): 23.9 ms against 17.0 ms with the row in shared memory;
  // phase 2 is bound by instruction issue, and the pair divides and converts every element twice.
  static const int stream_env = getenv("UPSP_PHASE2_STREAM") ? atoi(getenv("UPSP_PHASE2_STREAM")) : -1;
  if (a.itrans16 != nullptr && a.row_list == nullptr && a.F % 8 == 0 && a.blk_log2 == 0 &&
      (stream_env == 1 || (stream_env != 0 && phase2_cluster(a.F, true) == 0))) {
    int rc = UPSP_OK;
    for (int r0 = 0; r0 < a.n_local && !rc; r0 += 65535) {      // grid.y limit: rows in slabs
      Phase2Args b = a;
      b.n_local = std::min(65535, a.n_local - r0);
      b.node0 = a.node0 + r0;
      b.itrans16 = a.itrans16 + (size_t)r0 * a.F;
      b.ptrans = a.ptrans + (size_t)r0 * a.F;
      b.rms = a.rms + r0;
      b.avgp = a.avgp + r0;
      b.gain = a.gain + r0;
      switch (a.ncoef) {
#define UPSP_P2S(K) case K: rc = launch_phase2_stream<K>(c, b, st, launches); break;
        UPSP_P2S(1) UPSP_P2S(2) UPSP_P2S(3) UPSP_P2S(4) UPSP_P2S(5) UPSP_P2S(6) UPSP_P2S(7) UPSP_P2S(8) UPSP_P2S(9)
#undef UPSP_P2S
        default: return fail(UPSP_ERR_INVALID, "detrend degree %d not in [0,%d]", a.ncoef - 1, UPSP_MAX_COEF - 1);
      }
    }
    return rc;
  }
  switch (a.ncoef) {
    case 1: return launch_phase2<1>(c, a, st, launches);
    case 2: return launch_phase2<2>(c, a, st, launches);
    case 3: return launch_phase2<3>(c, a, st, launches);
    case 4: return launch_phase2<4>(c, a, st, launches);
    case 5: return launch_phase2<5>(c, a, st, launches);
    case 6: return launch_phase2<6>(c, a, st, launches);
    case 7: return launch_phase2<7>(c, a, st, launches);
    case 8: return launch_phase2<8>(c, a, st, launches);
    case 9: return launch_phase2<9>(c, a, st, launches);
  }
  return fail(UPSP_ERR_INVALID, "detrend degree %d not in [0,%d]", a.ncoef - 1, UPSP_MAX_COEF - 1);
}

// cluster size for a row of F frames: smallest CL in {1,2,4,8} whose per-CTA segment leaves room
// for 2 CTAs per SM (<= 90 KB + 18.6 KB static each); 0 = does not fit even with 8 (two HBM passes).
// 16-bit rows that need a cluster anyway take twice that many CTAs of 256 threads (<= 45 KB each, four CTAs per SM:
// four rows' barrier phases overlap instead of two).  Measured on one GPU, phase 2 of 10^10 node-frames (gpurun_out/r2an,
// r2ao): 40 000-frame rows 19.5 ms as 2 x 512 threads, 18.4 as 4 x 256; 80 000: 20.6 as 4 x 512, 18.8 as 8 x 256;
// 160 000: 21.9 as 8 x 512, 21.4 as 8 x 256.  Rows that fit one CTA stay there: 20 000 frames 16.4 ms as 1 x 512, 17.4 as
// 2 x 256, 19.9 as 4 x 128.  UPSP_P2_CL / UPSP_P2_NT force a shape.
static int phase2_cluster(int F, bool in16) {
  static const int forced = getenv("UPSP_P2_CL") ? atoi(getenv("UPSP_P2_CL")) : 0;
  if ((forced == 1 || forced == 2 || forced == 4 || forced == 8) && F % (8 * forced) == 0 &&
      (size_t)(F / forced) * sizeof(float) <= 200 * 1024)
    return forced;
  const size_t seg_budget = 90 * 1024, seg_max = 200 * 1024;
  for (int cl = 1; cl <= 8; cl *= 2) {
    const size_t seg = (size_t)(((F + cl * 4 - 1) / (cl * 4)) * 4) * sizeof(float);
    if (seg <= seg_budget || (cl == 8 && seg <= seg_max)) {
      if (in16 && (cl == 2 || cl == 4) && F % (16 * cl) == 0) return 2 * cl;
      return cl;
    }
  }
  return 0;
}
static bool phase2_symmetric(const Phase2Args& a) {
  const int cl = phase2_cluster(a.F, a.itrans16 != nullptr);
  return a.fit_out == nullptr && cl > 0 && a.F % (8 * cl) == 0 &&
         (reinterpret_cast<uintptr_t>(a.itrans) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.ptrans) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(a.itrans16) & 15) == 0;
}

static void fill_basis(Phase2Args& a, int F, int degree) {
  a.F = F;
  a.ncoef = degree + 1;
  a.xa = 2.0f / (float)F;
  a.xb = (1.0f - (float)F) / (float)F;
  // the O(F * ncoef^2) long-double Gram matrix depends on (F, degree, symmetric) only: memoise it
  // (10 ms at F = 80 000, which would otherwise sit in every phase-2 call)
  struct Key { int F, nc, sym; double ginv[UPSP_MAX_COEF * UPSP_MAX_COEF]; };
  static thread_local std::vector<Key> cache;
  const int sym = phase2_symmetric(a) ? 1 : 0;
  for (const Key& k : cache)
    if (k.F == F && k.nc == a.ncoef && k.sym == sym) {
      memcpy(a.ginv, k.ginv, sizeof a.ginv);
      return;
    }
  cheb_ginv(F, a.ncoef, a.xa, a.xb, sym != 0, a.ginv);
  Key k;
  k.F = F;
  k.nc = a.ncoef;
  k.sym = sym;
  memcpy(k.ginv, a.ginv, sizeof a.ginv);
  if (cache.size() >= 8) cache.erase(cache.begin());
  cache.push_back(k);
}

extern "C" int upsp_gpu_phase2(upsp_gpu_ctx* c, const upsp_phase2_params* p, const float* steady,
                               const float* model_temp) {
  ENTER(c);
  REQUIRE(p && steady && model_temp, UPSP_ERR_INVALID, "null argument");
  REQUIRE(c->phase1_done && c->transposed, UPSP_ERR_STATE,
          "phase2 needs finish_phase1 and transpose first");
  REQUIRE(p->degree >= 0 && p->degree < UPSP_MAX_COEF, UPSP_ERR_INVALID, "degree %d", p->degree);
  CU(cudaMemcpyAsync(c->d_steady, steady, (size_t)c->N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_temp, model_temp, (size_t)c->N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  Phase2Args a{};
  a.itrans = c->d_itrans;
  a.ptrans = c->d_ptrans;
  a.n_local = c->N_local;
  a.node0 = c->n0;
  a.avg = c->d_avg;
  a.coverage = c->d_cov;
  a.steady = c->d_steady;
  a.temp = c->d_temp;
  memcpy(a.cal, p->paint_cal, sizeof a.cal);
  a.qbar = p->qbar;
  a.ps = p->ps;
  fill_basis(a, c->F, p->degree);
  a.rms = c->d_rms2;
  a.avgp = c->d_avg2;
  a.gain = c->d_gain2;
  a.fit_out = nullptr;
  CU(cudaEventRecord(c->ev_a, c->stream));
  const bool prof = c->sample_every > 0;
  KBEGIN(6);
  if (c->it16) {
    // 16-bit rows of the plain nodes, then the float rows of the side buffer (patched / unseen nodes)
    Phase2Args b = a;
    b.itrans16 = reinterpret_cast<const unsigned short*>(c->d_shared);
    b.other_idx = c->d_other_idx;
    if (c->blk_len > 0) {
      int lg = 0;
      while ((1 << lg) < c->blk_len) ++lg;
      b.blk_log2 = lg;
      b.blk_kb = c->blk_kb;
      b.blk_floc = c->F / c->R;
      b.blk_rows = c->N_local;
      b.blk_magic = (unsigned)((((unsigned long long)1) << 32) / (unsigned)b.blk_floc + 1);
    }
    TRY(dispatch_phase2(c, b, c->stream, &c->launches));
    if (c->n_other_local > 0) {
      Phase2Args o = a;
      o.itrans = reinterpret_cast<const float*>(c->d_shared + c->side_off[c->rank]);
      o.other_idx = c->d_other_idx;
      o.row_list = c->d_other_local;
      o.n_local = c->n_other_local;
      TRY(dispatch_phase2(c, o, c->stream, &c->launches));
    }
  } else {
    TRY(dispatch_phase2(c, a, c->stream, &c->launches));
  }
  KEND();
  if (c->N_local) {
    k_phase2_finals<<<cdiv(c->N_local, 256), 256, 0, c->stream>>>(
        c->d_rms2, c->d_avg2, c->d_gain2, c->N_local, (unsigned)c->F, c->d_rms2f, c->d_avg2f, c->d_gain2f);
    KCHECK(c);
  }
  CU(cudaEventRecord(c->ev_b, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaEventElapsedTime(&c->stage_ms[3], c->ev_a, c->ev_b));
  c->phase2_done = true;
  return UPSP_OK;
}

// ------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------
extern "C" int upsp_gpu_sync(upsp_gpu_ctx* c) {
  ENTER(c);
  CU(cudaStreamSynchronize(c->copy_stream));
  CU(cudaStreamSynchronize(c->stream));
  return UPSP_OK;
}

static int d2h(upsp_gpu_ctx* c, void* host, const void* dev, size_t bytes) {
  CU(cudaStreamSynchronize(c->stream));
  if (bytes) CU(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost));
  return UPSP_OK;
}

extern "C" int upsp_gpu_read_intensity(upsp_gpu_ctx* c, int off, int n, float* host) {
  ENTER(c);
  REQUIRE(off >= 0 && n >= 0 && off + n <= c->F_local && (host || n == 0), UPSP_ERR_INVALID, "bad range");
  REQUIRE(c->finalized && !c->fused, UPSP_ERR_STATE,
          "frame-major intensity is not materialised (create the context with keep_frame_major=1)");
  REQUIRE(!(c->phase2_done && !c->ptrans_owned), UPSP_ERR_STATE,
          "frame-major intensity was overwritten by pressure_transpose (pressure_aliases_intensity)");
  return d2h(c, host, c->d_intensity + (size_t)off * c->N, (size_t)n * c->N * sizeof(float));
}

// 16-bit row mode: rows [row0, row0 + nrows) x frames [f0, f0 + nf) of intensity_transpose as floats (the values the
// float rows would hold: integers are exact, the side buffer's rows are copied)
__global__ void __launch_bounds__(256)
k_itrans_rows_f32(const uint16_t* __restrict__ it16, const float* __restrict__ side, const int* __restrict__ other_idx,
                  int node0, int row0, int F, int f0, int nf, float* __restrict__ out, size_t out_pitch, int blk_len,
                  int blk_kb, int blk_floc, int blk_rows) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (f >= nf) return;
  const int oi = __ldg(other_idx + node0 + row0 + r);
  const size_t col = (size_t)f0 + f;
  float v;
  if (oi >= 0) {
    v = side[(size_t)oi * F + col];
  } else if (blk_len > 0) {      // batch-blocked rows: [source rank * blk_kb + local batch][blk_rows][blk_len]
    const int src = (int)(col / (size_t)blk_floc), o = (int)(col - (size_t)src * blk_floc);
    const size_t blk = (size_t)src * blk_kb + (size_t)(o / blk_len);
    v = (float)it16[(blk * blk_rows + (size_t)(row0 + r)) * blk_len + (size_t)(o % blk_len)];
  } else {
    v = (float)it16[(size_t)(row0 + r) * F + col];
  }
  out[(size_t)r * out_pitch + f] = v;
}

// widen rows [noff, noff+nn) x frames [foff, foff+nf) into the bounce buffer chunk by chunk and copy each chunk out
static int read_itrans16(upsp_gpu_ctx* c, int noff, int nn, int foff, int nf, float* host, size_t pitch, cudaStream_t st) {
  if (nn == 0 || nf == 0) return UPSP_OK;
  const size_t want = std::min((size_t)nn * nf, (size_t)64 << 20);         // <= 256 MB of floats
  if (c->bounce_floats < std::max(want, (size_t)nf)) {
    CU(cudaStreamSynchronize(st));
    cudaFree(c->d_bounce);
    c->d_bounce = nullptr;
    c->bounce_floats = std::max(want, (size_t)nf);
    TRY(dmalloc(&c->d_bounce, c->bounce_floats));
  }
  const int rows_per = (int)std::max<size_t>(1, std::min<size_t>(c->bounce_floats / nf, 65535));
  const uint16_t* it16 = reinterpret_cast<const uint16_t*>(c->d_shared);
  const float* side = reinterpret_cast<const float*>(c->d_shared + c->side_off[c->rank]);
  for (int r0 = 0; r0 < nn; r0 += rows_per) {
    const int nr = std::min(rows_per, nn - r0);
    k_itrans_rows_f32<<<dim3(cdiv(nf, 256), nr), 256, 0, st>>>(it16, side, c->d_other_idx, c->n0, noff + r0, c->F, foff, nf,
                                                               c->d_bounce, (size_t)nf, c->blk_len, c->blk_kb,
                                                               c->blk_len > 0 ? c->F / c->R : 0, c->N_local);
    KCHECK(c);
    CU(cudaMemcpy2DAsync(host + (size_t)r0 * pitch, pitch * sizeof(float), c->d_bounce, (size_t)nf * sizeof(float),
                         (size_t)nf * sizeof(float), (size_t)nr, cudaMemcpyDeviceToHost, st));
  }
  return UPSP_OK;
}

extern "C" int upsp_gpu_read_intensity_transpose(upsp_gpu_ctx* c, int off, int n, float* host) {
  ENTER(c);
  REQUIRE(off >= 0 && n >= 0 && off + n <= c->N_local && (host || n == 0), UPSP_ERR_INVALID, "bad range");
  REQUIRE(c->transposed, UPSP_ERR_STATE, "transpose first");
  if (c->it16) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->d2h_stream));      // the bounce buffer is shared with the streamed block reads
    TRY(read_itrans16(c, off, n, 0, c->F, host, (size_t)c->F, c->d2h_stream));
    CU(cudaStreamSynchronize(c->d2h_stream));
    return UPSP_OK;
  }
  return d2h(c, host, c->d_itrans + (size_t)off * c->F, (size_t)n * c->F * sizeof(float));
}

extern "C" int upsp_gpu_read_intensity_transpose_block_async(upsp_gpu_ctx* c, int noff, int nn, int foff,
                                                             int nf, float* host, size_t pitch) {
  ENTER(c);
  REQUIRE(noff >= 0 && nn >= 0 && noff + nn <= c->N_local, UPSP_ERR_INVALID, "bad node range");
  REQUIRE(foff >= 0 && nf >= 0 && foff + nf <= c->F, UPSP_ERR_INVALID, "bad frame range");
  REQUIRE((host && pitch >= (size_t)nf) || nn == 0 || nf == 0, UPSP_ERR_INVALID, "bad host buffer / pitch");
  REQUIRE(c->finalized && c->fused, UPSP_ERR_STATE,
          "column-block reads need the fused projection (keep_frame_major = 0)");
  {
    const int lo = std::max(foff, c->f0), hi = std::min(foff + nf, c->f0 + c->F_local);   // own frames in the block
    REQUIRE(hi <= lo || hi - c->f0 <= c->frames_processed, UPSP_ERR_STATE,
            "frames [%d,%d) have not been submitted to process_frames yet", lo, hi);
  }
  if (nn == 0 || nf == 0) return UPSP_OK;
  CU(cudaStreamWaitEvent(c->d2h_stream, c->ev_proc, 0));
  if (c->it16) return read_itrans16(c, noff, nn, foff, nf, host, pitch, c->d2h_stream);
  CU(cudaMemcpy2DAsync(host, pitch * sizeof(float), c->d_itrans + (size_t)noff * c->F + foff,
                       (size_t)c->F * sizeof(float), (size_t)nf * sizeof(float), (size_t)nn,
                       cudaMemcpyDeviceToHost, c->d2h_stream));
  return UPSP_OK;
}

extern "C" int upsp_gpu_wait_pushes(upsp_gpu_ctx* c) {
  ENTER(c);
  CU(cudaStreamSynchronize(c->copy_stream));
  return UPSP_OK;
}

extern "C" int upsp_gpu_wait_reads(upsp_gpu_ctx* c) {
  ENTER(c);
  CU(cudaStreamSynchronize(c->d2h_stream));
  return UPSP_OK;
}

extern "C" int upsp_gpu_read_pressure_transpose(upsp_gpu_ctx* c, int off, int n, float* host) {
  ENTER(c);
  REQUIRE(off >= 0 && n >= 0 && off + n <= c->N_local && (host || n == 0), UPSP_ERR_INVALID, "bad range");
  REQUIRE(c->phase2_done, UPSP_ERR_STATE, "phase2 first");
  return d2h(c, host, c->d_ptrans + (size_t)off * c->F, (size_t)n * c->F * sizeof(float));
}

extern "C" int upsp_gpu_read_phase1_stats(upsp_gpu_ctx* c, float* avg, float* rms, float* cov) {
  ENTER(c);
  REQUIRE(c->phase1_done, UPSP_ERR_STATE, "finish_phase1 first");
  if (avg) TRY(d2h(c, avg, c->d_avg, (size_t)c->N * sizeof(float)));
  if (rms) TRY(d2h(c, rms, c->d_rms, (size_t)c->N * sizeof(float)));
  if (cov) TRY(d2h(c, cov, c->d_cov, (size_t)c->N * sizeof(float)));
  return UPSP_OK;
}

extern "C" int upsp_gpu_read_phase2_stats(upsp_gpu_ctx* c, float* rms, float* avg, float* gain) {
  ENTER(c);
  REQUIRE(c->phase2_done, UPSP_ERR_STATE, "phase2 first");
  if (rms) TRY(d2h(c, rms, c->d_rms2f, (size_t)c->N_local * sizeof(float)));
  if (avg) TRY(d2h(c, avg, c->d_avg2f, (size_t)c->N_local * sizeof(float)));
  if (gain) TRY(d2h(c, gain, c->d_gain2f, (size_t)c->N_local * sizeof(float)));
  return UPSP_OK;
}

extern "C" int upsp_gpu_read_warp_matrices(upsp_gpu_ctx* c, int cam, int off, int count, float* m6,
                                           float* rho, int* iters) {
  ENTER(c);
  CAM_CHECK(c, cam);
  Camera& k = c->cams[cam];
  REQUIRE(off >= 0 && count >= 0 && off + count <= c->F_local, UPSP_ERR_INVALID, "bad range");
  REQUIRE(k.d_m6, UPSP_ERR_STATE, "camera %d has no warp matrices", cam);
  if (m6) TRY(d2h(c, m6, k.d_m6 + (size_t)off * 6, (size_t)count * 6 * sizeof(float)));
  if (rho || iters) {
    REQUIRE(k.d_eccState, UPSP_ERR_STATE, "no ECC results (registration is not `pixel`)");
    std::vector<EccState> st(count);
    TRY(d2h(c, st.data(), k.d_eccState + off, (size_t)count * sizeof(EccState)));
    for (int f = 0; f < count; ++f) {
      if (rho) rho[f] = st[f].rho;
      if (iters) iters[f] = st[f].iters;
    }
  }
  return UPSP_OK;
}

extern "C" int upsp_gpu_stage_ms(upsp_gpu_ctx* c, int stage, float* ms) {
  ENTER(c);
  REQUIRE(stage >= 0 && stage < 4 && ms, UPSP_ERR_INVALID, "stage %d", stage);
  if (stage == 0 && c->stage_ms[0] < 0.0f) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventElapsedTime(&c->stage_ms[0], c->ev_pa, c->ev_pb));
  }
  *ms = c->stage_ms[stage];
  return UPSP_OK;
}

extern "C" int upsp_gpu_reset_timers(upsp_gpu_ctx* c) {
  ENTER(c);
  for (float& m : c->stage_ms) m = 0.0f;
  return UPSP_OK;
}

extern "C" int upsp_gpu_timer_start(upsp_gpu_ctx* c) {
  ENTER(c);
  CU(cudaEventRecord(c->ev_t0, c->stream));
  return UPSP_OK;
}

extern "C" int upsp_gpu_timer_stop(upsp_gpu_ctx* c, float* ms) {
  ENTER(c);
  REQUIRE(ms, UPSP_ERR_INVALID, "null argument");
  CU(cudaEventRecord(c->ev_t1, c->stream));
  CU(cudaEventSynchronize(c->ev_t1));
  CU(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
  return UPSP_OK;
}

extern "C" int upsp_gpu_reset_run(upsp_gpu_ctx* c) {
  ENTER(c);
  CU(cudaMemsetAsync(c->d_sum, 0, (size_t)c->N * sizeof(double), c->stream));
  CU(cudaMemsetAsync(c->d_sumsq, 0, (size_t)c->N * sizeof(double), c->stream));
  c->phase1_done = c->transposed = c->phase2_done = false;
  c->frames_processed = 0;
  c->kn = 0;
  c->batch_counter = 0;
  for (auto& r : c->proc_recs) r.count = 0;     // frame offsets start over: the records of the last run mean nothing now
  c->push_wait_all = true;                      // ... and the next push waits for whatever is still queued
  return UPSP_OK;
}

extern "C" int upsp_gpu_set_kernel_sampling(upsp_gpu_ctx* c, int every) {
  ENTER(c);
  REQUIRE(every >= 0, UPSP_ERR_INVALID, "sample_every %d", every);
  c->sample_every = every;
  return UPSP_OK;
}

extern "C" int upsp_gpu_kernel_ms(upsp_gpu_ctx* c, int cls, float* mean_ms, int* n_sampled) {
  ENTER(c);
  REQUIRE(mean_ms && n_sampled && cls >= 0 && cls <= 6, UPSP_ERR_INVALID, "bad argument");
  CU(cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  int n = 0;
  for (size_t i = 0; i < c->kn; ++i) {
    if (c->kcls[i] != cls) continue;
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, c->kev[2 * i], c->kev[2 * i + 1]));
    tot += ms;
    ++n;
  }
  *mean_ms = n ? (float)(tot / n) : 0.0f;
  *n_sampled = n;
  return UPSP_OK;
}

extern "C" int upsp_gpu_projection_mode(const upsp_gpu_ctx* c, int* mode) {
  REQUIRE(c && mode, UPSP_ERR_INVALID, "null argument");
  *mode = c->proj_mode;
  return UPSP_OK;
}

extern "C" int upsp_gpu_row_bytes(const upsp_gpu_ctx* c, int* bytes) {
  REQUIRE(c && bytes, UPSP_ERR_INVALID, "null argument");
  *bytes = c->it16 ? 2 : 4;
  return UPSP_OK;
}

extern "C" int upsp_gpu_timeline(upsp_gpu_ctx* c, int on) {
  ENTER(c);
  CU(cudaDeviceSynchronize());
  c->timeline = on != 0;
  c->kn = 0;
  return UPSP_OK;
}

extern "C" int upsp_gpu_timeline_read(upsp_gpu_ctx* c, float* rec, int max_records, int* n_records) {
  ENTER(c);
  REQUIRE(rec && n_records, UPSP_ERR_INVALID, "null argument");
  CU(cudaDeviceSynchronize());
  int n = 0;
  for (size_t i = 0; i < c->kn && n < max_records; ++i, ++n) {
    float a = 0.0f, b = 0.0f;
    CU(cudaEventElapsedTime(&a, c->kev[0], c->kev[2 * i]));
    CU(cudaEventElapsedTime(&b, c->kev[0], c->kev[2 * i + 1]));
    rec[3 * n] = (float)c->kcls[i];
    rec[3 * n + 1] = a;
    rec[3 * n + 2] = b;
  }
  *n_records = n;
  return UPSP_OK;
}

extern "C" int upsp_gpu_launch_count(const upsp_gpu_ctx* c, long long* n) {
  REQUIRE(c && n, UPSP_ERR_INVALID, "null argument");
  *n = c->launches;
  return UPSP_OK;
}

// ------------------------------------------------------------------------------------------
// multi-GPU wiring
// ------------------------------------------------------------------------------------------
extern "C" int upsp_gpu_ipc_export(upsp_gpu_ctx* c, void* handle) {
  ENTER(c);
  REQUIRE(handle, UPSP_ERR_INVALID, "null handle");
  REQUIRE(c->R > 1 && c->shared_vmm.fd >= 0, UPSP_ERR_STATE, "single-rank context has nothing to export");
  IpcWire w{};
  w.magic = 0x55505356u;
  w.pid = (int32_t)getpid();
  w.fd = c->shared_vmm.fd;
  w.device = c->cfg.device;
  w.size = c->shared_vmm.size;
  memcpy(handle, &w, sizeof w);
  return UPSP_OK;
}

extern "C" int upsp_gpu_ipc_import(upsp_gpu_ctx* c, const void* handles) {
  ENTER(c);
  REQUIRE(handles, UPSP_ERR_INVALID, "null handles");
  DriverApi& d = drv();
  REQUIRE(d.ok, UPSP_ERR_COMM, "CUDA VMM driver entry points unavailable");
  for (int r = 0; r < c->R; ++r) {
    if (r == c->rank) continue;
    IpcWire w;
    memcpy(&w, (const char*)handles + (size_t)r * UPSP_IPC_HANDLE_BYTES, sizeof w);
    REQUIRE(w.magic == 0x55505356u, UPSP_ERR_INVALID, "handle of rank %d is not a upsp_gpu handle", r);
    int fd = -1;
    if (w.pid == (int32_t)getpid()) {
      fd = dup(w.fd);
    } else {
      // duplicate the exporter's descriptor into this process (Linux >= 5.6)
      const int pidfd = (int)syscall(434 /* SYS_pidfd_open */, (pid_t)w.pid, 0u);
      REQUIRE(pidfd >= 0, UPSP_ERR_COMM, "pidfd_open(rank %d, pid %d) failed: %s", r, w.pid, strerror(errno));
      fd = (int)syscall(438 /* SYS_pidfd_getfd */, pidfd, w.fd, 0u);
      const int e = errno;
      close(pidfd);
      REQUIRE(fd >= 0, UPSP_ERR_COMM, "pidfd_getfd(rank %d) failed: %s (needs ptrace access to the peer process)",
              r, strerror(e));
    }
    VmmBlock& b = c->peer_vmm[r];
    b.size = (size_t)w.size;
    CUresult cr = d.MemImportFromShareableHandle(&b.handle, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    REQUIRE(cr == CUDA_SUCCESS, UPSP_ERR_COMM, "cuMemImportFromShareableHandle(rank %d) -> %d", r, (int)cr);
    TRY(vmm_map(b, c->cfg.device));
    c->peer_base[r] = reinterpret_cast<char*>(b.ptr);
    c->peer_is_ipc[r] = true;
  }
  c->peers_ready = true;
  return UPSP_OK;
}

extern "C" int upsp_gpu_connect_local(upsp_gpu_ctx** ctxs, int n) {
  REQUIRE(ctxs && n >= 1, UPSP_ERR_INVALID, "bad argument");
  for (int i = 0; i < n; ++i) {
    REQUIRE(ctxs[i] && ctxs[i]->R == n && ctxs[i]->rank == i, UPSP_ERR_INVALID,
            "context %d must be rank %d of %d", i, i, n);
  }
  for (int i = 0; i < n; ++i) {
    CU(cudaSetDevice(ctxs[i]->cfg.device));
    for (int j = 0; j < n; ++j) {
      if (i == j) continue;
      if (ctxs[i]->cfg.device != ctxs[j]->cfg.device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, ctxs[i]->cfg.device, ctxs[j]->cfg.device));
        REQUIRE(can, UPSP_ERR_COMM, "device %d cannot access device %d", ctxs[i]->cfg.device,
                ctxs[j]->cfg.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->cfg.device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else REQUIRE(e == cudaSuccess, UPSP_ERR_COMM, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        if (ctxs[j]->shared_vmm.handle) {   // VMM block: grant device i access to rank j's mapping
          CUmemAccessDesc acc{};
          acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
          acc.location.id = ctxs[i]->cfg.device;
          acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
          CUresult cr = drv().MemSetAccess(ctxs[j]->shared_vmm.ptr, ctxs[j]->shared_vmm.size, &acc, 1);
          REQUIRE(cr == CUDA_SUCCESS, UPSP_ERR_COMM, "cuMemSetAccess(device %d on rank %d's block) -> %d",
                  ctxs[i]->cfg.device, j, (int)cr);
        }
      }
      ctxs[i]->peer_base[j] = ctxs[j]->d_shared;
    }
    ctxs[i]->peers_ready = true;
  }
  return UPSP_OK;
}

extern "C" int upsp_gpu_set_exchange(upsp_gpu_ctx* c, int exchange) {
  ENTER(c);
  REQUIRE(exchange == UPSP_XCHG_PEER || exchange == UPSP_XCHG_NCCL, UPSP_ERR_INVALID, "exchange %d", exchange);
  REQUIRE(!c->finalized, UPSP_ERR_STATE, "set_exchange after the first process_frames");
  if (exchange == UPSP_XCHG_NCCL) REQUIRE(nccl().ok, UPSP_ERR_COMM, "libnccl.so.2 not found (UPSP_NCCL_LIB=path)");
  c->exchange = exchange;
  return UPSP_OK;
}
extern "C" int upsp_gpu_nccl_unique_id(void* id) {
  REQUIRE(id != nullptr, UPSP_ERR_INVALID, "null id");
  REQUIRE(nccl().ok, UPSP_ERR_COMM, "libnccl.so.2 not found (UPSP_NCCL_LIB=path)");
  NCCLCHK(nccl().GetUniqueId(reinterpret_cast<NcclApi::UniqueId*>(id)));
  return UPSP_OK;
}
extern "C" int upsp_gpu_nccl_init(upsp_gpu_ctx* c, const void* id) {
  ENTER(c);
  REQUIRE(id != nullptr, UPSP_ERR_INVALID, "null id");
  REQUIRE(nccl().ok, UPSP_ERR_COMM, "libnccl.so.2 not found (UPSP_NCCL_LIB=path)");
  REQUIRE(c->nccl_comm == nullptr, UPSP_ERR_STATE, "communicator already created");
  NcclApi::UniqueId uid;
  memcpy(&uid, id, sizeof uid);
  NCCLCHK(nccl().CommInitRank(&c->nccl_comm, c->R, uid, c->rank));
  return UPSP_OK;
}

// ------------------------------------------------------------------------------------------
// stand-alone operators
// ------------------------------------------------------------------------------------------
struct DevScope {
  int dev;
  std::vector<void*> bufs;
  explicit DevScope(int d) : dev(d) {}
  ~DevScope() {
    for (void* p : bufs) cudaFree(p);
  }
  template <typename T>
  int alloc(T** p, size_t n) {
    int rc = dmalloc(p, n);
    if (!rc) bufs.push_back((void*)*p);
    return rc;
  }
  template <typename T>
  int put(T** p, const T* h, size_t n) {
    TRY(alloc(p, n));
    if (n) CU(cudaMemcpy(*p, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return UPSP_OK;
  }
};
#define OP_ENTER(device)                                                                \
  {                                                                                     \
    int nd_ = upsp_gpu_device_count();                                                  \
    REQUIRE(nd_ > 0, UPSP_ERR_CUDA, "no usable CUDA device; libupsp_gpu has no CPU fallback"); \
    REQUIRE((device) >= 0 && (device) < nd_, UPSP_ERR_INVALID, "device %d of %d", device, nd_); \
    CU(cudaSetDevice(device));                                                          \
  }                                                                                     \
  DevScope S(device)

extern "C" int upsp_op_unpack(int device, const uint8_t* packed, int format, size_t npix,
                              const uint16_t* lut, uint16_t* out) {
  OP_ENTER(device);
  REQUIRE(packed && out && npix > 0, UPSP_ERR_INVALID, "null/empty argument");
  REQUIRE(format == UPSP_PIX_PACKED12 || format == UPSP_PIX_PACKED10, UPSP_ERR_INVALID, "format %d", format);
  REQUIRE(npix % (format == UPSP_PIX_PACKED12 ? 2 : 4) == 0, UPSP_ERR_INVALID, "pixel count %zu", npix);
  const size_t nbytes = frame_bytes_of(format, npix);
  uint8_t* d_in;
  uint16_t *d_out, *d_lut = nullptr;
  int *d_cnt, *d_pos;
  TRY(S.put(&d_in, packed, nbytes));
  TRY(S.alloc(&d_out, npix));
  TRY(S.alloc(&d_cnt, 1));
  TRY(S.alloc(&d_pos, UPSP_HOT_STORE));
  if (lut) TRY(S.put(&d_lut, lut, 1024));
  CU(cudaMemset(d_cnt, 0, sizeof(int)));
  if (format == UPSP_PIX_PACKED12)
    k_unpack12_scan<<<dim3(cdiv(cdiv(npix, 8), 256), 1), 256>>>(d_in, nbytes, d_out, npix, 0x7fffffff, d_cnt, d_pos, nullptr, 1, (int)npix);
  else
    k_unpack10_scan<<<dim3(cdiv(cdiv(npix, 4), 256), 1), 256>>>(d_in, nbytes, d_out, npix, d_lut, 0x7fffffff, d_cnt, d_pos, nullptr, 1, (int)npix);
  CU(cudaGetLastError());
  CU(cudaMemcpy(out, d_out, npix * 2, cudaMemcpyDeviceToHost));
  return UPSP_OK;
}

extern "C" int upsp_op_fix_hot_pixels(int device, uint16_t* frames, int nf, int rows, int cols, int* n_hot) {
  OP_ENTER(device);
  REQUIRE(frames && nf > 0 && rows > 0 && cols > 0, UPSP_ERR_INVALID, "null/empty argument");
  const size_t npix = (size_t)rows * cols;
  uint16_t *d_in, *d_out;
  int *d_cnt, *d_pos;
  TRY(S.put(&d_in, frames, npix * nf));
  TRY(S.alloc(&d_out, npix * nf));
  TRY(S.alloc(&d_cnt, nf));
  TRY(S.alloc(&d_pos, (size_t)nf * UPSP_HOT_STORE));
  CU(cudaMemset(d_cnt, 0, sizeof(int) * nf));
  k_copy16_scan<<<dim3(cdiv(cdiv(npix, 8), 256), nf), 256>>>(d_in, npix, d_out, npix, UPSP_HOT_THRESH, d_cnt, d_pos, nullptr, rows, cols);
  CU(cudaGetLastError());
  k_fix_hot<<<cdiv(nf, 32), 32>>>(d_out, npix, rows, cols, nf, d_cnt, d_pos, UPSP_HOT_MIN_CHANGE, UPSP_HOT_MAX);
  CU(cudaGetLastError());
  CU(cudaMemcpy(frames, d_out, npix * nf * 2, cudaMemcpyDeviceToHost));
  if (n_hot) {
    std::vector<int> cnt(nf);
    CU(cudaMemcpy(cnt.data(), d_cnt, sizeof(int) * nf, cudaMemcpyDeviceToHost));
    for (int f = 0; f < nf; ++f) n_hot[f] = cnt[f] > UPSP_HOT_MAX ? -1 : cnt[f];
  }
  return UPSP_OK;
}

extern "C" int upsp_op_warp_affine(int device, const uint16_t* src, int nf, int W, int H,
                                   const float* m6, int interp, uint16_t* dst) {
  OP_ENTER(device);
  REQUIRE(src && dst && m6 && nf > 0 && W > 0 && H > 0, UPSP_ERR_INVALID, "null/empty argument");
  REQUIRE(interp == 0 || interp == 1, UPSP_ERR_INVALID, "interp %d", interp);
  const size_t npix = (size_t)W * H;
  uint16_t *d_src, *d_dst;
  float* d_m;
  int* d_tab;
  TRY(S.put(&d_src, src, npix * nf));
  TRY(S.alloc(&d_dst, npix * nf));
  TRY(S.put(&d_m, m6, (size_t)nf * 6));
  TRY(S.alloc(&d_tab, (size_t)nf * (2 * W + 2 * H)));
  k_warp_tables<<<dim3(cdiv(std::max(W, H), 256), nf), 256>>>(d_m, nf, W, H, interp, d_tab);
  CU(cudaGetLastError());
  k_warp_affine8_u16<<<dim3(cdiv(W, 1024), H, nf), 128>>>(d_src, d_dst, W, H, d_tab, interp, -1);
  CU(cudaGetLastError());
  CU(cudaMemcpy(dst, d_dst, npix * nf * 2, cudaMemcpyDeviceToHost));
  return UPSP_OK;
}

extern "C" int upsp_op_project_frames(int device, const int32_t* rowptr, const int32_t* col,
                                      const float* val, int n_rows, const float* frames, int nf,
                                      size_t npix, float* out) {
  OP_ENTER(device);
  REQUIRE(rowptr && frames && out && n_rows > 0 && nf > 0 && npix > 0, UPSP_ERR_INVALID, "null/empty argument");
  const int nnz = rowptr[n_rows];
  for (int i = 0; i < nnz; ++i)
    REQUIRE(col[i] >= 0 && (size_t)col[i] < npix, UPSP_ERR_INVALID, "column %d outside the frame", col[i]);
  int *d_rp, *d_col;
  float *d_val, *d_fr, *d_out;
  TRY(S.put(&d_rp, rowptr, (size_t)n_rows + 1));
  TRY(S.put(&d_col, col, (size_t)nnz));
  TRY(S.put(&d_val, val, (size_t)nnz));
  TRY(S.put(&d_fr, frames, npix * nf));
  TRY(S.alloc(&d_out, (size_t)n_rows * nf));
  k_project_f32<<<dim3(cdiv(n_rows, 256), nf), 256>>>(d_rp, d_col, d_val, n_rows, d_fr, npix, d_out);
  CU(cudaGetLastError());
  CU(cudaMemcpy(out, d_out, (size_t)n_rows * nf * sizeof(float), cudaMemcpyDeviceToHost));
  return UPSP_OK;
}

extern "C" int upsp_op_transpose(int device, const float* src, int x_extent, int y_extent, float* dst) {
  OP_ENTER(device);
  REQUIRE(src && dst && x_extent > 0 && y_extent > 0, UPSP_ERR_INVALID, "null/empty argument");
  float *d_src, *d_dst;
  const size_t n = (size_t)x_extent * y_extent;
  TRY(S.put(&d_src, src, n));
  TRY(S.alloc(&d_dst, n));
  XposeArgs a{};
  a.src = d_src;
  a.rows = y_extent;
  a.cols = x_extent;
  a.n_ranks = 1;
  a.f_total = y_extent;
  a.col0 = 0;
  a.dst[0] = d_dst;
  a.node_start[0] = 0;
  a.node_start[1] = x_extent;
  k_transpose_a2a<<<dim3(cdiv(x_extent, XT), cdiv(y_extent, XT)), 256>>>(a);
  CU(cudaGetLastError());
  CU(cudaMemcpy(dst, d_dst, n * sizeof(float), cudaMemcpyDeviceToHost));
  return UPSP_OK;
}

extern "C" int upsp_op_polyfit_detrend(int device, const float* data, int n_pts, int n_frames,
                                       int degree, float* fit) {
  OP_ENTER(device);
  REQUIRE(data && fit && n_pts > 0 && n_frames > 0, UPSP_ERR_INVALID, "null/empty argument");
  REQUIRE(degree >= 0 && degree < UPSP_MAX_COEF, UPSP_ERR_INVALID, "degree %d", degree);
  float *d_data, *d_fit;
  const size_t n = (size_t)n_pts * n_frames;
  TRY(S.put(&d_data, data, n));
  TRY(S.alloc(&d_fit, n));
  Phase2Args a{};
  a.itrans = d_data;
  a.ptrans = d_fit;
  a.fit_out = d_fit;
  a.n_local = n_pts;
  fill_basis(a, n_frames, degree);
  long long l = 0;
  TRY(dispatch_phase2(nullptr, a, 0, &l));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(fit, d_fit, n * sizeof(float), cudaMemcpyDeviceToHost));
  return UPSP_OK;
}

// ------------------------------------------------------------------------------------------
// phase-0 product: projection matrix (kernels_setup.cuh)
// ------------------------------------------------------------------------------------------
static void rodrigues_f64(const double r[3], double R[9]) {   // cv::Rodrigues(rvec -> R)
  const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (theta < 2.220446049250313e-16) {
    for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, itheta = 1.0 / theta;
  const double x = r[0] * itheta, y = r[1] * itheta, z = r[2] * itheta;
  const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
  const double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int k = 0; k < 9; ++k) R[k] = c * I[k] + c1 * rrt[k] + s * rx[k];
}

static int fill_setup_cam(const upsp_camera_model* cam, SetupCam& sc) {
  REQUIRE(cam && cam->width > 0 && cam->height > 0, UPSP_ERR_INVALID, "camera model");
  rodrigues_f64(cam->rvec, sc.R);
  for (int i = 0; i < 3; ++i) sc.t[i] = cam->tvec[i];
  sc.fx = cam->fx; sc.fy = cam->fy; sc.cx = cam->cx; sc.cy = cam->cy;
  for (int i = 0; i < 8; ++i) sc.k[i] = cam->dist[i];
  for (int i = 0; i < 3; ++i)   // CameraCal::get_cam_center: -R^T t (double), narrowed to float
    sc.orig[i] = (float)(-(sc.R[0 + i] * cam->tvec[0] + sc.R[3 + i] * cam->tvec[1] + sc.R[6 + i] * cam->tvec[2]));
  sc.width = cam->width;
  sc.height = cam->height;
  return UPSP_OK;
}

extern "C" int upsp_op_project_points(int device, const upsp_camera_model* cam, const float* xyz, int n, float* uv) {
  OP_ENTER(device);
  REQUIRE(xyz && uv && n > 0, UPSP_ERR_INVALID, "null/empty argument");
  SetupCam sc{};
  TRY(fill_setup_cam(cam, sc));
  float *d_xyz, *d_uv;
  TRY(S.put(&d_xyz, xyz, (size_t)3 * n));
  TRY(S.alloc(&d_uv, (size_t)2 * n));
  k_project_points<<<cdiv(n, 128), 128>>>(sc, d_xyz, n, d_uv);
  CU(cudaGetLastError());
  CU(cudaMemcpy(uv, d_uv, (size_t)2 * n * sizeof(float), cudaMemcpyDeviceToHost));
  return UPSP_OK;
}

extern "C" int upsp_op_create_projection(int device, const upsp_camera_model* cam, const float* xyz,
                                         const float* normals, const uint8_t* is_datanode, int n_nodes,
                                         const int32_t* tri_nodes, int n_tris, float oblique_thresh,
                                         int32_t* code, float* uv) {
  OP_ENTER(device);
  REQUIRE(xyz && normals && is_datanode && tri_nodes && code && uv && n_nodes > 0 && n_tris > 0, UPSP_ERR_INVALID,
          "null/empty argument");
  for (size_t i = 0; i < (size_t)3 * n_tris; ++i)
    REQUIRE(tri_nodes[i] >= 0 && tri_nodes[i] < n_nodes, UPSP_ERR_INVALID, "triangle %zu references node %d", i / 3, tri_nodes[i]);
  SetupCam sc{};
  TRY(fill_setup_cam(cam, sc));
  // ---- direction-space grid (host, once): pinhole coordinates of every vertex seen from the ray origin
  std::vector<double> xn(n_nodes), yn(n_nodes);
  std::vector<uint8_t> front(n_nodes);
  double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
  for (int i = 0; i < n_nodes; ++i) {
    const double w[3] = {(double)xyz[3 * i] - (double)sc.orig[0], (double)xyz[3 * i + 1] - (double)sc.orig[1],
                         (double)xyz[3 * i + 2] - (double)sc.orig[2]};
    const double qx = sc.R[0] * w[0] + sc.R[1] * w[1] + sc.R[2] * w[2], qy = sc.R[3] * w[0] + sc.R[4] * w[1] + sc.R[5] * w[2];
    const double qz = sc.R[6] * w[0] + sc.R[7] * w[1] + sc.R[8] * w[2];
    const double len = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    front[i] = qz > 1e-3 * len;             // well in front of the camera plane (within ~89.9 deg of the axis)
    if (!front[i]) continue;
    xn[i] = qx / qz;
    yn[i] = qy / qz;
    lo[0] = std::min(lo[0], xn[i]); hi[0] = std::max(hi[0], xn[i]);
    lo[1] = std::min(lo[1], yn[i]); hi[1] = std::max(hi[1], yn[i]);
  }
  SetupGrid grid{};
  std::vector<int> cell_start(1, 0), cell_tris, always;
  if (lo[0] <= hi[0]) {
    const int G = std::max(8, std::min(2048, (int)sqrt((double)n_tris / 2.0)));
    const double ex = std::max(hi[0] - lo[0], 1e-9), ey = std::max(hi[1] - lo[1], 1e-9);
    grid.gx = grid.gy = G;
    // the gridded range is padded by one cell on every side so that jittered rays stay inside
    const double cwx = ex / (G - 2), cwy = ey / (G - 2);
    grid.x0 = lo[0] - cwx;
    grid.y0 = lo[1] - cwy;
    grid.inv_cell_x = 1.0 / cwx;
    grid.inv_cell_y = 1.0 / cwy;
    auto cell_range = [&](double a, double b, double o, double inv, int g, int& c0, int& c1) {
      // margin of a quarter cell: orders of magnitude above the float error of the edge tests and
      // above the 1e-4 jitter of the retry rays
      c0 = std::max(0, (int)floor((a - o) * inv - 0.25));
      c1 = std::min(g - 1, (int)floor((b - o) * inv + 0.25));
    };
    std::vector<int> count((size_t)G * G + 1, 0);
    for (int pass = 0; pass < 2; ++pass) {
      for (int k = 0; k < n_tris; ++k) {
        const int a = tri_nodes[3 * k], b = tri_nodes[3 * k + 1], c = tri_nodes[3 * k + 2];
        if (!(front[a] && front[b] && front[c])) {
          if (pass == 0) always.push_back(k);
          continue;
        }
        int x0, x1, y0, y1;
        cell_range(std::min({xn[a], xn[b], xn[c]}), std::max({xn[a], xn[b], xn[c]}), grid.x0, grid.inv_cell_x, G, x0, x1);
        cell_range(std::min({yn[a], yn[b], yn[c]}), std::max({yn[a], yn[b], yn[c]}), grid.y0, grid.inv_cell_y, G, y0, y1);
        for (int y = y0; y <= y1; ++y)
          for (int x = x0; x <= x1; ++x) {
            if (pass == 0) count[(size_t)y * G + x + 1]++;
            else cell_tris[(size_t)cell_start[(size_t)y * G + x] + count[(size_t)y * G + x]++] = k;
          }
      }
      if (pass == 0) {
        cell_start.assign((size_t)G * G + 1, 0);
        for (size_t i = 0; i < (size_t)G * G; ++i) cell_start[i + 1] = cell_start[i] + count[i + 1];
        cell_tris.resize((size_t)cell_start.back());
        std::fill(count.begin(), count.end(), 0);
      }
    }
  }
  float *d_xyz, *d_nrm, *d_uv;
  uint8_t* d_isd;
  int *d_tri, *d_code, *d_cs = nullptr, *d_ct = nullptr, *d_al = nullptr;
  TRY(S.put(&d_xyz, xyz, (size_t)3 * n_nodes));
  TRY(S.put(&d_nrm, normals, (size_t)3 * n_nodes));
  TRY(S.put(&d_isd, is_datanode, (size_t)n_nodes));
  TRY(S.put(&d_tri, tri_nodes, (size_t)3 * n_tris));
  TRY(S.alloc(&d_code, (size_t)n_nodes));
  TRY(S.alloc(&d_uv, (size_t)2 * n_nodes));
  if (grid.gx) {
    TRY(S.put(&d_cs, cell_start.data(), cell_start.size()));
    if (!cell_tris.empty()) TRY(S.put(&d_ct, cell_tris.data(), cell_tris.size()));
    else TRY(S.alloc(&d_ct, 1));
  }
  if (!always.empty()) TRY(S.put(&d_al, always.data(), always.size()));
  grid.cell_start = d_cs;
  grid.cell_tris = d_ct;
  grid.always = d_al;
  grid.n_always = (int)always.size();
  k_create_projection<<<cdiv(n_nodes, 128), 128>>>(sc, grid, d_xyz, d_nrm, d_isd, n_nodes, d_tri, n_tris, oblique_thresh,
                                                   d_code, d_uv);
  CU(cudaGetLastError());
  CU(cudaMemcpy(code, d_code, (size_t)n_nodes * sizeof(int), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(uv, d_uv, (size_t)2 * n_nodes * sizeof(float), cudaMemcpyDeviceToHost));
  return UPSP_OK;
}
