// upsp_b200.hpp -- C++17 host side over the C ABI of libupsp_gpu.so (include/upsp_gpu.h).
//
// The reference's hot path is C++ (cpp/exec/psp_process.cpp phase1()/phase2() calling
// cpp/lib); this header mirrors the pieces of that interface the GPU library replaces, with the
// reference's names and argument meaning, so a maintainer can swap call sites one by one:
//
//   reference                                              here
//   ------------------------------------------------------ ------------------------------------
//   apportion()                    psp_process.cpp:611     upsp_b200::apportion
//   unpack_12bit / unpack_10bit    PSPVideo.cpp:111-150    upsp_b200::unpack_12bit / unpack_10bit
//   upsp::fix_hot_pixels           cv_extras.cpp:230       upsp_b200::fix_hot_pixels
//   cv::warpAffine in register_pixel registration.cpp:69   upsp_b200::warp_affine
//   upsp::project_frame            projection.ipp:884      upsp_b200::project_frame
//   local_transpose                psp_process.cpp:647     upsp_b200::local_transpose
//   TransPolyFitter<float>         filtering.ipp:13-76     upsp_b200::TransPolyFitter
//   phase1() frame loop + global_transpose + phase2() node loop   upsp_b200::FrameChain
//   pwrite_full / write_block / output_files  :524-540, :627-639, :958-963   upsp_b200::FlatFiles
//
// Errors: the reference asserts / throws (cv::Exception, std::invalid_argument) and psp_process
// turns that into exit code 1; every wrapper here throws upsp_b200::Error carrying the library's
// status code and message.  No CPU fallback: everything computes on the GPU.
#pragma once

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/upsp_gpu.h"

namespace upsp_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void check(int rc) {
  if (rc != UPSP_OK) throw Error(rc, upsp_gpu_last_error());
}

/** Divide [0, value) into nBins contiguous slices, the first `value % nBins` one longer
 *  (psp_process.cpp:611-624). */
inline void apportion(int value, int nBins, int* start, int* extent) {
  unsigned long block = value / nBins, rem = value - block * nBins, next = 0;
  for (unsigned long b = 0; b < (unsigned long)nBins; ++b) {
    start[b] = (int)next;
    extent[b] = (int)(block + (b < rem));
    next += extent[b];
  }
}

// ---- library seams (stand-alone operators; host buffers in and out) -----------------------
inline void unpack_12bit(const uint8_t* packed, uint16_t* unpacked, size_t n_pixels, int device = 0) {
  check(upsp_op_unpack(device, packed, UPSP_PIX_PACKED12, n_pixels, nullptr, unpacked));
}
inline void unpack_10bit(const uint8_t* packed, uint16_t* unpacked, size_t n_pixels,
                         const uint16_t* lut1024 = nullptr, int device = 0) {
  check(upsp_op_unpack(device, packed, UPSP_PIX_PACKED10, n_pixels, lut1024, unpacked));
}
/** In place on n_frames frames of rows x cols; n_hot[f] = hot pixels found, -1 = too many. */
inline void fix_hot_pixels(uint16_t* frames, int n_frames, int rows, int cols, int* n_hot = nullptr,
                           int device = 0) {
  check(upsp_op_fix_hot_pixels(device, frames, n_frames, rows, cols, n_hot));
}
/** cv::warpAffine(src, warp_matrix, size, interp | WARP_INVERSE_MAP); m6 = n_frames 2x3 maps. */
inline void warp_affine(const uint16_t* src, int n_frames, int width, int height, const float* m6,
                        int interp, uint16_t* dst, int device = 0) {
  check(upsp_op_warp_affine(device, src, n_frames, width, height, m6, interp, dst));
}
/** output[f][row] = smat(row,:) . frame f  (Eigen row-major CSR). */
inline void project_frame(const int32_t* rowptr, const int32_t* col, const float* val, int n_rows,
                          const float* frames, int n_frames, size_t n_pixels, float* output,
                          int device = 0) {
  check(upsp_op_project_frames(device, rowptr, col, val, n_rows, frames, n_frames, n_pixels, output));
}
/** dst[x][y] = src[y][x]; x_extent = contiguous extent of src. */
inline void local_transpose(const float* src, int x_extent, int y_extent, float* dst, int device = 0) {
  check(upsp_op_transpose(device, src, x_extent, y_extent, dst));
}

/** upsp::TransPolyFitter<float>(n_frames, degree, n_pts).eval_fit(data, n_pts, 0). */
class TransPolyFitter {
 public:
  TransPolyFitter(unsigned n_frames, unsigned degree, unsigned /*n_pts*/, int device = 0)
      : n_frames_(n_frames), degree_(degree), device_(device) {}
  /** data: [n_pts][n_frames] (node-major); returns the fitted values, same layout. */
  std::vector<float> eval_fit(const float* data, unsigned n_pts, unsigned /*curr_pt*/ = 0) const {
    std::vector<float> out((size_t)n_pts * n_frames_);
    check(upsp_op_polyfit_detrend(device_, data, (int)n_pts, (int)n_frames_, (int)degree_, out.data()));
    return out;
  }

 private:
  unsigned n_frames_, degree_;
  int device_;
};

// ---- the frame chain ---------------------------------------------------------------------
/** RAII wrapper of one rank's context; call order == phase1() / global_transpose() / phase2(). */
class FrameChain {
 public:
  explicit FrameChain(const upsp_gpu_config& cfg) { check(upsp_gpu_create(&cfg, &ctx_)); }
  ~FrameChain() { upsp_gpu_destroy(ctx_); }
  FrameChain(const FrameChain&) = delete;
  FrameChain& operator=(const FrameChain&) = delete;
  upsp_gpu_ctx* raw() { return ctx_; }

  void slices(int& first_frame, int& n_frames, int& first_node, int& n_nodes) const {
    check(upsp_gpu_get_slices(ctx_, &first_frame, &n_frames, &first_node, &n_nodes));
  }
  void set_camera(int cam, int width, int height) { check(upsp_gpu_set_camera(ctx_, cam, width, height)); }
  void set_projection(int cam, const int32_t* rowptr, const int32_t* col, const float* val) {
    check(upsp_gpu_set_projection(ctx_, cam, rowptr, col, val));
  }
  void set_overlap_remap(const int32_t* src_index) { check(upsp_gpu_set_overlap_remap(ctx_, src_index)); }
  void set_options(int registration, int interp, int patcher, bool hot_pixel_fix = true) {
    check(upsp_gpu_set_options(ctx_, registration, interp, patcher, hot_pixel_fix));
  }
  /* deck @options filter / filter_size: kind 0 none, 1 gaussian, 2 box */
  void set_filter(int kind, int ksize) { check(upsp_gpu_set_filter(ctx_, kind, ksize)); }
  void set_patches(int cam, int n_clusters, const int32_t* boff, const uint32_t* bx, const uint32_t* by,
                   const int32_t* ioff, const uint32_t* ix, const uint32_t* iy) {
    check(upsp_gpu_set_patches(ctx_, cam, n_clusters, boff, bx, by, ioff, ix, iy));
  }
  void set_reference_frame(int cam, const uint16_t* first_frame) {
    check(upsp_gpu_set_reference_frame(ctx_, cam, first_frame));
  }
  void set_warp_matrices(int cam, int local_offset, int count, const float* m6) {
    check(upsp_gpu_set_warp_matrices(ctx_, cam, local_offset, count, m6));
  }
  /* 10 -> 12-bit table of packed 10-bit cines (CineReader.cpp:409-425) */
  void set_unpack_lut(const uint16_t* lut1024) { check(upsp_gpu_set_unpack_lut(ctx_, lut1024)); }
  void push_frames(int cam, const void* frames, int format, int local_offset, int count) {
    check(upsp_gpu_push_frames(ctx_, cam, frames, format, local_offset, count));
  }
  void wait_pushes() { check(upsp_gpu_wait_pushes(ctx_)); }     // the pushed host buffers may be refilled; processing goes on
  void process_frames(int local_offset, int count) { check(upsp_gpu_process_frames(ctx_, local_offset, count)); }
  void finish_phase1() { check(upsp_gpu_finish_phase1(ctx_)); }
  void global_transpose() { check(upsp_gpu_transpose(ctx_)); }
  void phase2(const upsp_phase2_params& p, const float* steady, const float* model_temp) {
    check(upsp_gpu_phase2(ctx_, &p, steady, model_temp));
  }
  void sync() { check(upsp_gpu_sync(ctx_)); }
  void read_intensity_transpose(int node_off, int n, float* host) {
    check(upsp_gpu_read_intensity_transpose(ctx_, node_off, n, host));
  }
  void read_pressure_transpose(int node_off, int n, float* host) {
    check(upsp_gpu_read_pressure_transpose(ctx_, node_off, n, host));
  }
  void read_phase1_stats(float* avg, float* rms, float* coverage) {
    check(upsp_gpu_read_phase1_stats(ctx_, avg, rms, coverage));
  }
  void read_phase2_stats(float* rms, float* avg, float* gain) {
    check(upsp_gpu_read_phase2_stats(ctx_, rms, avg, gain));
  }

 private:
  upsp_gpu_ctx* ctx_ = nullptr;
};

// ---- flat-file outputs --------------------------------------------------------------------
/** The reference's flat output files (psp_process.cpp:524-540): created by rank 0, written with
 *  pwrite at byte offsets, little-endian f32.  [N x F] files take each rank's node slice at
 *  offset start_node * F * 4 (write_block :958-963); [N] vectors are written whole by rank 0. */
class FlatFiles {
 public:
  FlatFiles(const std::string& dir, bool create) : dir_(dir), create_(create) {}
  ~FlatFiles() {
    for (auto& kv : fds_) close(kv.second);
  }
  /** Like pwrite, except it keeps going after a partial write (psp_process.cpp:627-639). */
  static void pwrite_full(int fd, const void* buf, size_t nbytes, off_t file_offset) {
    const unsigned char* src = static_cast<const unsigned char*>(buf);
    while (nbytes > 0) {
      const ssize_t w = pwrite(fd, src, nbytes, file_offset);
      if (w <= 0) throw Error(UPSP_ERR_INVALID, std::string("pwrite_full: ") + strerror(errno));
      nbytes -= (size_t)w;
      file_offset += w;
      src += w;
    }
  }
  void write_vector(const std::string& name, const float* v, size_t n) {
    pwrite_full(fd(name), v, n * sizeof(float), 0);
  }
  /** rows [start_node, start_node + n_nodes) of an [N x n_frames] node-major file. */
  void write_block(const std::string& name, const float* rows, size_t start_node, size_t n_nodes,
                   size_t n_frames) {
    pwrite_full(fd(name), rows, n_nodes * n_frames * sizeof(float),
                (off_t)(start_node * n_frames * sizeof(float)));
  }

 private:
  int fd(const std::string& name) {
    auto it = fds_.find(name);
    if (it != fds_.end()) return it->second;
    const std::string path = dir_ + "/" + name;
    const int f = open(path.c_str(), create_ ? (O_WRONLY | O_CREAT | O_TRUNC) : O_WRONLY, 0644);
    if (f < 0) throw Error(UPSP_ERR_INVALID, "cannot open " + path + ": " + strerror(errno));
    fds_[name] = f;
    return f;
  }
  std::string dir_;
  bool create_;
  std::map<std::string, int> fds_;
};

}  // namespace upsp_b200
