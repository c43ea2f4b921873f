// projection_weights_probe -- runs host/projection_weights.hpp on raw arrays (tests/test_projection_weights.py).
//   projection_weights_probe DIR n_cams n_nodes best|average
// DIR holds xyz.f32 [n][3], nrm.f32 [n][3], centers.f64 [n_cams][3], cam<c>.rowptr/.col/.val; the scaled
// values are written back to cam<c>.val.out and the skipped node list to skipped.u32.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "projection_weights.hpp"

template <typename T>
static std::vector<T> rd(const std::string& p) {
  std::ifstream f(p, std::ios::binary | std::ios::ate);
  if (!f) throw std::invalid_argument("Cannot open '" + p + "'");
  std::vector<T> v((size_t)f.tellg() / sizeof(T));
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
  return v;
}
template <typename T>
static void wr(const std::string& p, const std::vector<T>& v) {
  std::ofstream f(p, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}

int main(int argc, char** argv) {
  if (argc < 5) {
    std::cerr << "usage: projection_weights_probe DIR n_cams n_nodes best|average\n";
    return 1;
  }
  try {
    using namespace upsp_b200;
    const std::string d = argv[1];
    const int nc = std::atoi(argv[2]), n = std::atoi(argv[3]);
    const auto xyz = rd<float>(d + "/xyz.f32"), nrm = rd<float>(d + "/nrm.f32");
    const auto cen = rd<double>(d + "/centers.f64");
    if ((int)xyz.size() != 3 * n || (int)nrm.size() != 3 * n || (int)cen.size() != 3 * nc) throw std::invalid_argument("array sizes");
    std::vector<CsrMatrix> projs((size_t)nc);
    std::vector<std::array<double, 3>> centers((size_t)nc);
    for (int c = 0; c < nc; ++c) {
      const std::string b = d + "/cam" + std::to_string(c);
      projs[(size_t)c].rowptr = rd<int32_t>(b + ".rowptr");
      projs[(size_t)c].col = rd<int32_t>(b + ".col");
      projs[(size_t)c].val = rd<float>(b + ".val");
      centers[(size_t)c] = {cen[3 * (size_t)c], cen[3 * (size_t)c + 1], cen[3 * (size_t)c + 2]};
    }
    adjust_projection_for_weights(xyz.data(), nrm.data(), centers, projs,
                                  std::string(argv[4]) == "best" ? OverlapType::BestView : OverlapType::AverageViews);
    for (int c = 0; c < nc; ++c) wr(d + "/cam" + std::to_string(c) + ".val.out", projs[(size_t)c].val);
    std::vector<unsigned> skipped;
    identify_skipped_nodes(projs, skipped);
    wr(d + "/skipped.u32", skipped);
  } catch (const std::exception& e) {
    std::cerr << "projection_weights_probe: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
