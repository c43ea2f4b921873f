// patch_geometry_probe -- runs host/patch_geometry.hpp on a target list and prints the resulting
// cluster pixel lists (tests/test_patch_geometry.py compares them with an independent restatement).
//   patch_geometry_probe targets.txt width height boundary_thickness buffer_thickness [ref.u16 thresh offset]
// targets.txt: one "u v diameter" per line.  Output: "cluster i n_targets", then "b x y" / "i x y" lines.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "patch_geometry.hpp"

int main(int argc, char** argv) {
  if (argc < 6) {
    std::cerr << "usage: patch_geometry_probe targets.txt width height boundary_thickness buffer_thickness [ref.u16 thresh offset]\n";
    return 1;
  }
  using namespace upsp_b200;
  std::ifstream f(argv[1]);
  if (!f) {
    std::cerr << "Cannot open '" << argv[1] << "'\n";
    return 1;
  }
  std::vector<Target> targs;
  Target t;
  while (f >> t.u >> t.v >> t.diameter) targs.push_back(t);
  const int W = std::atoi(argv[2]), H = std::atoi(argv[3]);
  const unsigned bt = (unsigned)std::atoi(argv[4]), bf = (unsigned)std::atoi(argv[5]);
  std::vector<std::vector<Target>> clusters;
  cluster_points(targs, clusters, (int)(bt + bf));                 // psp_process.cpp:2127-2128
  PatchClusters pc(clusters, W, H, bt, bf);
  if (argc >= 9) {
    std::ifstream r(argv[6], std::ios::binary);
    std::vector<uint16_t> ref((size_t)W * H);
    r.read(reinterpret_cast<char*>(ref.data()), (std::streamsize)(ref.size() * 2));
    if (!r) {
      std::cerr << "Cannot read reference frame '" << argv[6] << "'\n";
      return 1;
    }
    pc.threshold_bounds(ref.data(), (unsigned)std::atoi(argv[7]), (unsigned)std::atoi(argv[8]));
  }
  for (size_t i = 0; i < clusters.size(); ++i) {
    std::printf("cluster %zu %zu\n", i, clusters[i].size());
    for (size_t j = 0; j < pc.bounds_x[i].size(); ++j) std::printf("b %u %u\n", pc.bounds_x[i][j], pc.bounds_y[i][j]);
    for (size_t j = 0; j < pc.internal_x[i].size(); ++j) std::printf("i %u %u\n", pc.internal_x[i][j], pc.internal_y[i][j]);
  }
  return 0;
}
