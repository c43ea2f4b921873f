/*
 * upsp_gpu.h -- C ABI of libupsp_gpu.so: the B200 (sm_100a) implementation of
 * psp_process's per-frame data-parallel chain.
 *
 * The reference (nasa/upsp-processing) has no plugin/FFI layer: the path is inlined in
 * phase1()/phase2() of cpp/exec/psp_process.cpp.  This header is the seam a maintainer
 * binds in place of the OpenMP loops at psp_process.cpp:1743-1851 (frame loop) and
 * :2452-2507 (node loop) and of global_transpose() (:707-771).  Each entry point cites
 * the reference interface it replaces.  See INTEGRATION.md for the C++ stub.
 *
 * Conventions: every function returns 0 (UPSP_OK) or a non-zero upsp_status and records
 * a message retrievable with upsp_gpu_last_error() (thread local).  The caller owns all
 * host memory; the library owns all device memory.  One context per GPU / per rank.  A
 * context is thread-compatible (external locking), not thread-safe.  There is no CPU
 * fallback: creation fails if no CUDA device is usable.
 */
#ifndef UPSP_GPU_H_
#define UPSP_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define UPSP_API __attribute__((visibility("default")))
#else
#define UPSP_API
#endif

typedef struct upsp_gpu_ctx upsp_gpu_ctx;

typedef enum {
  UPSP_OK = 0,
  UPSP_ERR_INVALID = 1, /* bad argument                                  */
  UPSP_ERR_STATE = 2,   /* call out of order (e.g. phase2 before transpose) */
  UPSP_ERR_CUDA = 3,    /* CUDA runtime error                             */
  UPSP_ERR_NOMEM = 4,   /* device or host allocation failed               */
  UPSP_ERR_COMM = 5,    /* multi-GPU wiring (IPC / NCCL) failed           */
  UPSP_ERR_NUMERIC = 6  /* ECC did not converge (cv::Exception in the reference) */
} upsp_status;

/* pixel container of pushed frames: PSPVideo bit depths (cpp/lib/PSPVideo.cpp:111-150) */
typedef enum { UPSP_PIX_U16 = 0, UPSP_PIX_PACKED12 = 1, UPSP_PIX_PACKED10 = 2 } upsp_pixel_format;
/* @options registration (cpp/lib/upsp_inputs.cpp; psp_process.cpp:1291-1295).
 * UPSP_REG_GIVEN: warp matrices supplied by the caller (upsp_gpu_set_warp_matrices). */
typedef enum { UPSP_REG_NONE = 0, UPSP_REG_PIXEL = 1, UPSP_REG_GIVEN = 2 } upsp_registration;
/* @options pixel_interpolation -> cv::INTER_NEAREST / cv::INTER_LINEAR (psp_process.cpp:1781-1788) */
typedef enum { UPSP_INTERP_NEAREST = 0, UPSP_INTERP_LINEAR = 1 } upsp_interp;
/* @options target_patcher (psp_process.cpp:1797) */
typedef enum { UPSP_PATCH_NONE = 0, UPSP_PATCH_POLYNOMIAL = 1 } upsp_patcher;
/* how upsp_gpu_transpose moves data between ranks */
typedef enum { UPSP_XCHG_PEER = 0 /* fused transpose + peer stores over NVLink */,
               UPSP_XCHG_NCCL = 1 /* local transpose + ncclSend/ncclRecv all-to-all */ } upsp_exchange;

typedef struct {
  int device;         /* CUDA device ordinal                                             */
  int n_cams;         /* ifile.cameras                                                    */
  int n_nodes;        /* msize = model.size()                                             */
  int n_frames_total; /* number_frames                                                    */
  int rank, n_ranks;  /* my_mpi_rank, num_mpi_ranks: frame / node slices by apportion()   */
  int frame_capacity; /* device input slots per camera; 0 => this rank's frame count
                         (whole slice resident).  Frame `o` lives in slot o % capacity.  */
  int batch_frames;   /* frames per internal launch batch; 0 => default (256)             */
  int pressure_aliases_intensity; /* 1: pressure_transpose reuses the frame-major
                         intensity storage when that buffer exists (saves one F x N buffer);
                         0: separate buffer                                              */
  int keep_frame_major; /* 0 (default): projections with <= 1 entry per row (the reference's
                         case) are written straight to node-major rows by the fused
                         register+project+transpose kernel; the frame-major buffer is never
                         materialised and upsp_gpu_read_intensity is unavailable.
                         1: materialise [F_local x N] and transpose separately.          */
} upsp_gpu_config;

/* Phase2Settings + TunnelConditions + PaintCalibration scalars (psp_process.cpp:1094-1104,
 * :2273-2310; cpp/include/non_cv_upsp.h:18-40) */
typedef struct {
  float paint_cal[6]; /* a..f of PaintCalibration::get_gain (non_cv_upsp.cpp:66-68) */
  float qbar;         /* tcond.qbar (psf) */
  float ps;           /* tcond.ps  (psf) */
  int degree;         /* sett.degree (6) */
} upsp_phase2_params;

/* ---- lifecycle ------------------------------------------------------------------- */
UPSP_API const char* upsp_gpu_last_error(void);
UPSP_API int upsp_gpu_device_count(void);
/* replaces allocate_global_data() psp_process.cpp:795-822 + apportion() :1520-1529 */
UPSP_API int upsp_gpu_create(const upsp_gpu_config* cfg, upsp_gpu_ctx** out);
UPSP_API int upsp_gpu_destroy(upsp_gpu_ctx* ctx);
/* frame / node slice of this rank: rank_start_frame/rank_num_frames/rank_start_node/
 * rank_num_nodes (psp_process.cpp:488-498) */
UPSP_API int upsp_gpu_get_slices(const upsp_gpu_ctx* ctx, int* first_frame, int* n_frames,
                                 int* first_node, int* n_nodes);

/* ---- setup (phase-0 products, replicated on every rank) ------------------------------ */
/* frame size of camera `cam` (elems.cals[c].size()) */
UPSP_API int upsp_gpu_set_camera(upsp_gpu_ctx* ctx, int cam, int width, int height);
/* elems.projs[cam]: Eigen::SparseMatrix<float,RowMajor> [n_nodes x width*height] in CSR
 * (psp_process.cpp:1604-1640; values already weighted by adjust_projection_for_weights).
 * rowptr[n_nodes+1], col/val[rowptr[n_nodes]].  Empty rows in every camera become the
 * `skipped` list of identify_skipped_nodes (projection.ipp:858-880). */
UPSP_API int upsp_gpu_set_projection(upsp_gpu_ctx* ctx, int cam, const int32_t* rowptr,
                                     const int32_t* col, const float* val);
/* static form of P3DModel_::adjust_solution (P3DModel.ipp:144-157): out[n] = in[src[n]].
 * NULL (or never called) == TriModel's no-op (Model.h:105). */
UPSP_API int upsp_gpu_set_overlap_remap(upsp_gpu_ctx* ctx, const int32_t* src_index);
/* deck @options (psp_process.cpp:1772-1807).  hot_pixel_fix: the reference always runs
 * upsp::fix_hot_pixels (thresh 4064, min_change 512, max_hot 5). */
UPSP_API int upsp_gpu_set_options(upsp_gpu_ctx* ctx, int registration, int interp, int patcher,
                                  int hot_pixel_fix);
/* deck @options filter / filter_size (psp_process.cpp:1802-1807): kind 0 none, 1 gaussian
 * (cv::GaussianBlur(img, img, Size(k,k), 0)), 2 box (cv::blur(img, img, Size(k,k))); k odd.
 * Gaussian: any odd size up to 31 on the CV_16U image (no patcher; OpenCV's fixed-point taps), 3 / 5 / 7 on the CV_32F
 * image the patcher returns.  Size 1 is the identity.  A filter makes the chain
 * materialise the registered image (no fused register+project kernel). */
UPSP_API int upsp_gpu_set_filter(upsp_gpu_ctx* ctx, int kind, int ksize);
/* PatchClusters geometry of camera `cam` (patches.h:74-90: bounds_x/y, internal_x/y per
 * cluster, after threshold_bounds).  Offsets are CSR-style [n_clusters+1]. */
UPSP_API int upsp_gpu_set_patches(upsp_gpu_ctx* ctx, int cam, int n_clusters,
                                  const int32_t* bounds_off, const uint32_t* bounds_x,
                                  const uint32_t* bounds_y, const int32_t* internal_off,
                                  const uint32_t* internal_x, const uint32_t* internal_y);
/* 10-bit cine -> 12-bit table (the caller owns CINE2_LUT, CineReader.cpp:23-87); NULL = none */
UPSP_API int upsp_gpu_set_unpack_lut(upsp_gpu_ctx* ctx, const uint16_t* lut1024);
/* elems.first_frames[cam] (raw first frame, u16): the ECC template of register_pixel
 * (psp_process.cpp:1790, :2057-2058) */
UPSP_API int upsp_gpu_set_reference_frame(upsp_gpu_ctx* ctx, int cam, const uint16_t* frame);
/* UPSP_REG_GIVEN: warp_matrix (2x3 f32, row-major) of local frames [offset, offset+count) */
UPSP_API int upsp_gpu_set_warp_matrices(upsp_gpu_ctx* ctx, int cam, int local_offset, int count,
                                        const float* m6);

/* ---- phase 1 ---------------------------------------------------------------------- */
/* replaces __async_read_ahead() filling input_frames[c][offset] (psp_process.cpp:867-908):
 * async H2D of `count` frames of camera `cam` starting at local frame `local_offset`. */
UPSP_API int upsp_gpu_push_frames(upsp_gpu_ctx* ctx, int cam, const void* host_frames, int format,
                                  int local_offset, int count);
/* Blocks until every push_frames issued so far has read its host buffer (the reader may refill it: the hand-over of
 * input_frames slots between __async_read_ahead and the frame loop, psp_process.cpp:897-908).  Does not wait for
 * process_frames; with an input ring smaller than the local slice it waits for the slots being overwritten to be consumed. */
UPSP_API int upsp_gpu_wait_pushes(upsp_gpu_ctx* ctx);
/* replaces the OpenMP frame loop psp_process.cpp:1753-1843 for local frames
 * [local_offset, local_offset+count): hot-pixel fix, register, patch, project, camera
 * sum, NaN fill, sum / sum-of-squares, overlap remap, row store.  Asynchronous. */
UPSP_API int upsp_gpu_process_frames(upsp_gpu_ctx* ctx, int local_offset, int count);
/* replaces the critical-section merge + MPI_Reduce x2 + finals + MPI_Bcast
 * (psp_process.cpp:1846-1872, :1930-1940, :2019-2023) */
UPSP_API int upsp_gpu_finish_phase1(upsp_gpu_ctx* ctx);
/* replaces global_transpose() psp_process.cpp:707-771 (local_transpose :647-689 +
 * Isend/Recv all-to-all + reassembly) */
UPSP_API int upsp_gpu_transpose(upsp_gpu_ctx* ctx);

/* ---- phase 2 ---------------------------------------------------------------------- */
/* replaces the OpenMP node loop psp_process.cpp:2452-2507 (+ TransPolyFitter::eval_fit
 * filtering.ipp:48-76, PaintCalibration::get_gain non_cv_upsp.cpp:66-68).
 * steady[n_nodes], model_temp[n_nodes] are indexed by global node. */
UPSP_API int upsp_gpu_phase2(upsp_gpu_ctx* ctx, const upsp_phase2_params* prm, const float* steady,
                             const float* model_temp);

/* ---- results (all synchronise the context's stream first) --------------------------- */
UPSP_API int upsp_gpu_sync(upsp_gpu_ctx* ctx);
/* frame-major intensity rows (the reference's ptr_intensity_data) */
UPSP_API int upsp_gpu_read_intensity(upsp_gpu_ctx* ctx, int local_frame_off, int n_frames,
                                     float* host);
/* node-major rows of this rank's node slice: `intensity_transpose`, `pressure_transpose`
 * flat files (psp_process.cpp:524-540, write_block :958-963) */
UPSP_API int upsp_gpu_read_intensity_transpose(upsp_gpu_ctx* ctx, int local_node_off, int n_nodes,
                                               float* host);
UPSP_API int upsp_gpu_read_pressure_transpose(upsp_gpu_ctx* ctx, int local_node_off, int n_nodes,
                                              float* host);
/* Streaming output (the reference's write-behind thread, psp_process.cpp:977-1007): asynchronous
 * D2H of a COLUMN block of this rank's intensity_transpose -- rows [local_node_off, +n_nodes),
 * GLOBAL frames [frame_off, +n_frames) -- into host[i * host_pitch + j] (host_pitch in floats;
 * pinned memory for real overlap).  Columns of this rank's own frames must already have been
 * submitted with process_frames; columns written by peer ranks are the caller's to order (read
 * them after the post-transpose barrier).  Ordered after every process_frames
 * call issued so far and executed on a dedicated copy stream, so it overlaps later pushes and
 * processing (PCIe is full duplex).  Needs the fused projection (keep_frame_major = 0), where a
 * node-major column is final as soon as its frames are processed.  upsp_gpu_wait_reads blocks
 * until all such reads have landed. */
UPSP_API int upsp_gpu_read_intensity_transpose_block_async(upsp_gpu_ctx* ctx, int local_node_off,
                                                           int n_nodes, int frame_off,
                                                           int n_frames, float* host,
                                                           size_t host_pitch);
UPSP_API int upsp_gpu_wait_reads(upsp_gpu_ctx* ctx);
/* sol_avg_final, sol_rms_final, coverage: [n_nodes] each; any pointer may be NULL */
UPSP_API int upsp_gpu_read_phase1_stats(upsp_gpu_ctx* ctx, float* avg, float* rms, float* coverage);
/* rms_final, avg_final, gain_final of this rank's node slice: [n_local_nodes] each */
UPSP_API int upsp_gpu_read_phase2_stats(upsp_gpu_ctx* ctx, float* rms, float* avg, float* gain);
/* registration results of local frames: m6[count][6], rho[count], iters[count] (any NULL) */
UPSP_API int upsp_gpu_read_warp_matrices(upsp_gpu_ctx* ctx, int cam, int local_offset, int count,
                                         float* m6, float* rho, int* iters);
/* device time (ms, CUDA events on the context stream) of the last call of each stage:
 * 0 process_frames (accumulated since create/reset), 1 finish_phase1, 2 transpose, 3 phase2 */
UPSP_API int upsp_gpu_stage_ms(upsp_gpu_ctx* ctx, int stage, float* ms);
UPSP_API int upsp_gpu_reset_timers(upsp_gpu_ctx* ctx);
/* bracket an arbitrary sequence of calls with CUDA events on the context's stream
 * (the stream every kernel of this context is launched on) */
UPSP_API int upsp_gpu_timer_start(upsp_gpu_ctx* ctx);
UPSP_API int upsp_gpu_timer_stop(upsp_gpu_ctx* ctx, float* ms);
/* Per-kernel device times: every `sample_every`-th internal batch (and every transpose /
 * phase-2 launch) is bracketed with CUDA events on the launching stream.  kernel_class:
 * 0 decode(+scan), 1 frame prep, 2 warp (unfused mode), 3 patch, 4 projection (fused: +transpose
 * +exchange), 5 transpose, 6 phase 2.  Returns the mean launch duration and the number of
 * sampled launches since the last upsp_gpu_reset_run / create.  A sampled batch is taken out of
 * the two-stream pipeline (its decode / patch run on the main stream, the next batch's front end
 * waits for its projection) so that the durations are those of the kernels alone. */
UPSP_API int upsp_gpu_set_kernel_sampling(upsp_gpu_ctx* ctx, int sample_every);
UPSP_API int upsp_gpu_kernel_ms(upsp_gpu_ctx* ctx, int kernel_class, float* mean_ms, int* n_sampled);
/* start a new run on the same context (same setup, same device buffers): zeroes the
 * sum / sum-sq accumulators and the phase flags; pushed frames stay resident */
UPSP_API int upsp_gpu_reset_run(upsp_gpu_ctx* ctx);
/* number of kernels this context has launched since create */
UPSP_API int upsp_gpu_launch_count(const upsp_gpu_ctx* ctx, long long* n);
/* which projection kernel family the context settled on at its first batch (-1 before it):
 * 0 = gather kernels (k_project_fused4 / k_project_fused / k_project_ell1 / k_project_csr),
 * 1 = TMA-staged boxes cut from the decoded u16 frames (k_project_tma<0>),
 * 2 = TMA-staged boxes cut from the packed 12-bit frames, no decode pass (k_project_tma<1> + k_hot_scan12).
 * Same results bit for bit; bench.py labels its per-kernel figures with it. */
UPSP_API int upsp_gpu_projection_mode(const upsp_gpu_ctx* ctx, int* mode);
/* Debug timeline: with `on` != 0 every kernel of the following batches is bracketed with CUDA events on its
 * own stream WITHOUT taking the batch out of the two-stream pipeline; upsp_gpu_timeline_read returns up to
 * `max_records` records {kernel_class, start_ms, end_ms} (ms since the first recorded event) and the count. */
/* bytes per stored value of the node-major intensity rows in device memory: 4 (float), or 2 once the context has
 * chosen 16-bit integer rows (one camera, unit projection values; decided at the first batch).  Reporting only: the
 * readers always deliver floats. */
UPSP_API int upsp_gpu_row_bytes(const upsp_gpu_ctx* ctx, int* bytes);
UPSP_API int upsp_gpu_timeline(upsp_gpu_ctx* ctx, int on);
UPSP_API int upsp_gpu_timeline_read(upsp_gpu_ctx* ctx, float* records3, int max_records, int* n_records);

/* ---- multi-GPU wiring (one process per GPU; bytes are exchanged by the host, e.g. with
 *      torch.distributed / MPI_Allgather) ------------------------------------------------ */
#define UPSP_IPC_HANDLE_BYTES 64
#define UPSP_NCCL_ID_BYTES 128
/* export this rank's intensity_transpose buffer; import everyone's (n_ranks * 64 bytes,
 * rank order).  After import, UPSP_XCHG_PEER transposes store directly into peer HBM. */
UPSP_API int upsp_gpu_ipc_export(upsp_gpu_ctx* ctx, void* handle);
UPSP_API int upsp_gpu_ipc_import(upsp_gpu_ctx* ctx, const void* handles);
/* UPSP_XCHG_NCCL: the reference's own structure on NCCL (replaces MPI_Isend / MPI_Recv of global_transpose,
 * cpp/exec/psp_process.cpp:707-771, and the MPI_Allreduce of the sums, :1866-1872): phase 1 keeps the frame-major
 * intermediate, upsp_gpu_transpose packs node-major blocks per destination rank (chunks of 1024 frames), moves them with
 * grouped ncclSend / ncclRecv and reassembles them; upsp_gpu_finish_phase1 uses ncclAllReduce.  No peer mappings are
 * needed (no ipc_export / ipc_import).  libnccl.so.2 is bound at run time (UPSP_NCCL_LIB=path overrides the search).
 * It is the measured comparison of the default UPSP_XCHG_PEER (exchange fused into the projection kernel), not the
 * fast path: DESIGN.md section 5.  Call set_exchange before the first process_frames, nccl_init on every rank with the
 * id that rank 0 got from nccl_unique_id. */
UPSP_API int upsp_gpu_nccl_unique_id(void* id128);
UPSP_API int upsp_gpu_nccl_init(upsp_gpu_ctx* ctx, const void* id128);
UPSP_API int upsp_gpu_set_exchange(upsp_gpu_ctx* ctx, int exchange);
/* single-process multi-GPU (tests, one host thread per context): wire contexts created on
 * different devices of the same process directly, no IPC */
UPSP_API int upsp_gpu_connect_local(upsp_gpu_ctx** ctxs, int n);

/* ---- stand-alone operators (host buffers in / out; the library seams of SURVEY 8b) ----- */
/* upsp::unpack_12bit / unpack_10bit (PSPVideo.cpp:111-150) */
UPSP_API int upsp_op_unpack(int device, const uint8_t* packed, int format, size_t n_pixels,
                            const uint16_t* lut1024, uint16_t* out);
/* upsp::fix_hot_pixels (cv_extras.cpp:230-272) on n_frames frames; n_hot[f] = count or -1 */
UPSP_API int upsp_op_fix_hot_pixels(int device, uint16_t* frames, int n_frames, int rows, int cols,
                                    int* n_hot);
/* cv::warpAffine(src, M, size, interp|WARP_INVERSE_MAP) (registration.cpp:69-72), batched */
UPSP_API int upsp_op_warp_affine(int device, const uint16_t* src, int n_frames, int width,
                                 int height, const float* m6, int interp, uint16_t* dst);
/* upsp::project_frame (projection.ipp:884-908) on n_frames f32 frames -> out[n_frames][n_rows] */
UPSP_API int upsp_op_project_frames(int device, const int32_t* rowptr, const int32_t* col,
                                    const float* val, int n_rows, const float* frames,
                                    int n_frames, size_t n_pixels, float* out);
/* local_transpose (psp_process.cpp:647-689): dst[x][y] = src[y][x] */
UPSP_API int upsp_op_transpose(int device, const float* src, int x_extent, int y_extent, float* dst);
/* TransPolyFitter<float>(n_frames, degree, n_pts).eval_fit(data, n_pts, 0)
 * (filtering.ipp:13-76): data/fit are [n_pts][n_frames] */
UPSP_API int upsp_op_polyfit_detrend(int device, const float* data, int n_pts, int n_frames,
                                     int degree, float* fit);

/* ---- phase-0 product on the GPU: the pixel-to-node projection matrix.
 * create_projection_mat (cpp/exec/psp_process.cpp:168-355): every data node is projected with the
 * camera calibration (cv::projectPoints, CameraCal.ipp:227-239), kept if it lands inside the frame,
 * if the nearest triangle hit by the ray from the camera centre (rt::BVH::intersect,
 * pspRT.cpp:359-430; six jittered retries) contains the node, and if the angle between the node
 * normal and the ray exceeds oblique_thresh; its row of the matrix is then the single entry
 * (nearest pixel, 1.0).
 *   xyz, normals [n_nodes][3] (Model::get_position / get_n), is_datanode [n_nodes],
 *   tri_nodes [n_tris][3] node indices of every triangle (the reference's triNodes),
 *   code [n_nodes]: pixel index y*width + x of the node's entry, or -1 (empty row),
 *   uv [2*n_nodes]: normalised image coordinates as psp_process writes them (0 for empty rows). */
typedef struct {
  double rvec[3], tvec[3];        /* CameraCal rvec_/tvec_ */
  double fx, fy, cx, cy;          /* cameraMatrix_ */
  double dist[8];                 /* distCoeffs_: k1 k2 p1 p2 k3 k4 k5 k6 (4, 5 or 8 used; rest 0) */
  int width, height;
} upsp_camera_model;
UPSP_API int upsp_op_create_projection(int device, const upsp_camera_model* cam, const float* xyz,
                                       const float* normals, const uint8_t* is_datanode, int n_nodes,
                                       const int32_t* tri_nodes, int n_tris, float oblique_thresh,
                                       int32_t* code, float* uv);
/* cv::projectPoints for n float points (CameraCal::map_points_to_image): uv [n][2] */
UPSP_API int upsp_op_project_points(int device, const upsp_camera_model* cam, const float* xyz, int n, float* uv);

#ifdef __cplusplus
}
#endif
#endif /* UPSP_GPU_H_ */
