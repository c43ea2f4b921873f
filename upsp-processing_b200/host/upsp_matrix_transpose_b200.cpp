// upsp_matrix_transpose_b200 -- the reference's stand-alone pressure <-> pressure_transpose tool
// (cpp/exec/upsp_matrix_transpose.cpp:276-430) on one GPU.  Same positional arguments, same output
// file names:
//   upsp_matrix_transpose_b200 msize number_frames flag_transpose inputDataFile outputDataFolder [device]
//     flag_transpose = 0: input is `pressure`           [number_frames x msize] -> writes pressure_transpose
//     flag_transpose = 1: input is `pressure_transpose` [msize x number_frames] -> writes pressure
// The reference distributes row blocks over MPI ranks and exchanges tiles; here row blocks of the
// input (as many rows as fit a 1 GiB staging buffer) go through upsp_op_transpose (k_transpose_a2a,
// 64x64 shared-memory tiles) and every transposed block is written with pwrite at its place in the
// output, i.e. the block structure of general_global_transpose (:147-212) with one rank.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/upsp_gpu.h"

static void pread_full(int fd, void* buf, size_t n, off_t off) {       // counterpart of pwrite_full (:627-639)
  char* p = static_cast<char*>(buf);
  while (n) {
    const ssize_t r = pread(fd, p, n, off);
    if (r <= 0) throw std::runtime_error("short read of the input matrix");
    p += r; n -= (size_t)r; off += r;
  }
}
static void pwrite_full(int fd, const void* buf, size_t n, off_t off) {
  const char* p = static_cast<const char*>(buf);
  while (n) {
    const ssize_t r = pwrite(fd, p, n, off);
    if (r <= 0) throw std::runtime_error("write failed");
    p += r; n -= (size_t)r; off += r;
  }
}

int main(int argc, char** argv) {
  if (argc < 6) {
    std::cerr << "usage: upsp_matrix_transpose_b200 msize number_frames flag_transpose inputDataFile outputDataFolder [device]\n";
    return 1;
  }
  try {
    const size_t msize = std::stoul(argv[1]), number_frames = std::stoul(argv[2]);
    const int flag_transpose = std::stoi(argv[3]);
    const std::string in_name = argv[4], out_dir = argv[5];
    const int device = argc > 6 ? std::atoi(argv[6]) : 0;
    std::cout << "\nInput arguments:\n\nmsize = " << msize << "\nnumbeR_frames = " << number_frames
              << "\nflag_transpose = " << flag_transpose << "\ninputDataFileName = " << in_name
              << "\noutputDataFolderName = " << out_dir << std::endl;
    // input is [rows x cols], output [cols x rows]
    const size_t rows = flag_transpose == 0 ? number_frames : msize;
    const size_t cols = flag_transpose == 0 ? msize : number_frames;
    const std::string out_name = out_dir + (flag_transpose == 0 ? "/pressure_transpose" : "/pressure");
    const int in_fd = open(in_name.c_str(), O_RDONLY);
    if (in_fd < 0) throw std::invalid_argument("Cannot open '" + in_name + "'");
    struct stat st{};
    fstat(in_fd, &st);
    if ((size_t)st.st_size != rows * cols * sizeof(float))
      throw std::invalid_argument("'" + in_name + "' holds " + std::to_string(st.st_size) + " bytes, expected msize*number_frames*4");
    unlink(out_name.c_str());
    const int out_fd = open(out_name.c_str(), O_WRONLY | O_CREAT, 0644);
    if (out_fd < 0) throw std::invalid_argument("Cannot create '" + out_name + "'");
    std::cout << (flag_transpose == 0 ? "Reading pressure data ..." : "Reading pressure_transpose data ...") << std::endl;
    // staging buffer: 1 GiB of floats (UPSP_XPOSE_BLOCK_FLOATS overrides: tests force several blocks)
    const size_t stage = getenv("UPSP_XPOSE_BLOCK_FLOATS") ? std::stoul(getenv("UPSP_XPOSE_BLOCK_FLOATS")) : ((size_t)1 << 28);
    const size_t block_rows = std::max<size_t>(1, std::min(rows, stage / cols));
    std::vector<float> src(block_rows * cols), dst(block_rows * cols);
    std::cout << "Construct the transpose" << std::endl;
    for (size_t r0 = 0; r0 < rows; r0 += block_rows) {
      const size_t nr = std::min(block_rows, rows - r0);
      pread_full(in_fd, src.data(), nr * cols * sizeof(float), (off_t)(r0 * cols * sizeof(float)));
      if (upsp_op_transpose(device, src.data(), (int)cols, (int)nr, dst.data()) != UPSP_OK)
        throw std::runtime_error(upsp_gpu_last_error());
      if (nr == rows) {
        pwrite_full(out_fd, dst.data(), nr * cols * sizeof(float), 0);
      } else {          // dst is [cols x nr]: row c lands at columns [r0, r0+nr) of output row c
        for (size_t c = 0; c < cols; ++c)
          pwrite_full(out_fd, dst.data() + c * nr, nr * sizeof(float), (off_t)((c * rows + r0) * sizeof(float)));
      }
    }
    std::cout << "transpose complete\n" << (flag_transpose == 0 ? "Writing pressure_transpose data ..." : "Writing pressure data ...")
              << "\nData writing complete" << std::endl;
    close(in_fd);
    close(out_fd);
  } catch (const std::exception& e) {
    std::cerr << "upsp_matrix_transpose_b200: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
