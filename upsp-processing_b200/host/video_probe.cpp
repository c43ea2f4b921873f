// video_probe -- prints the properties of a .cine / .mraw file and the CRC-32 of the stored bytes
// of selected frames (what upsp_gpu_push_frames would receive); optionally dumps the frames.
// Used by tests/test_video_readers.py to pin host/video_readers.hpp against the reference's readers.
//   video_probe FILE [first_frame=1] [count=all] [dump.bin]
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "video_readers.hpp"

static uint32_t crc32(const uint8_t* p, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::cerr << "usage: video_probe FILE [first_frame] [count] [dump.bin]\n";
    return 1;
  }
  try {
    auto v = upsp_b200::open_video(argv[1]);
    const auto& p = v->properties();
    const unsigned first = argc > 2 ? (unsigned)std::atoi(argv[2]) : 1;
    const unsigned count = argc > 3 ? (unsigned)std::atoi(argv[3]) : p.num_frames - first + 1;
    std::printf("width %u\nheight %u\nbit_depth %u\nnum_frames %u\nframe_rate %.6g\naperture %.6g\nexposure %.6g\n"
                "pixel_format %d\nframe_bytes %zu\nhas_lut %d\n",
                p.width, p.height, p.bit_depth, p.num_frames, p.frame_rate, (double)p.aperture, (double)p.exposure,
                v->pixel_format(), v->frame_bytes(), v->unpack_lut() ? 1 : 0);
    std::vector<uint8_t> buf(v->frame_bytes());
    FILE* dump = argc > 4 ? std::fopen(argv[4], "wb") : nullptr;
    for (unsigned n = first; n < first + count; ++n) {
      v->read_packed(n, buf.data());
      std::printf("crc %u %u\n", n, crc32(buf.data(), buf.size()));
      if (dump) std::fwrite(buf.data(), 1, buf.size(), dump);
    }
    if (dump) std::fclose(dump);
    if (v->unpack_lut()) std::printf("lut_crc %u\n", crc32(reinterpret_cast<const uint8_t*>(v->unpack_lut()), 2048));
  } catch (const std::exception& e) {
    std::cerr << "video_probe: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
