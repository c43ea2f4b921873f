// grid_probe -- reads a .tri grid with host/grid_readers.hpp, prints its sizes and dumps
// xyz | normals | tris (0-based) as raw little-endian arrays for tests/test_grid_readers.py; or reads
// an unformatted plot3d grid and re-writes it in single or double precision.
//   grid_probe FILE.tri [dump_prefix [intersect]]
//   grid_probe FILE.x sp|dp [OUT.x]
#include <cstdio>
#include <iostream>

#include "grid_readers.hpp"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::cerr << "usage: grid_probe FILE.tri [dump_prefix]\n";
    return 1;
  }
  try {
    const std::string name = argv[1];
    if (name.size() > 2 && name.compare(name.size() - 2, 2, ".x") == 0) {
      const bool dp = argc > 2 && std::string(argv[2]) == "dp";
      auto report = [](const auto& g) {
        std::printf("n_zones %u\nn_points %zu\n", g.num_zones(), g.size());
        for (unsigned z = 0; z < g.num_zones(); ++z)
          std::printf("zone %u %u %u %u\n", z, g.grid_size[z][0], g.grid_size[z][1], g.grid_size[z][2]);
      };
      if (dp) {
        upsp_b200::StructuredGrid<double> g;
        upsp_b200::read_plot3d_grid_file(name, g);
        report(g);
        if (argc > 3) upsp_b200::write_plot3d_grid_file(argv[3], g);
      } else {
        upsp_b200::StructuredGrid<float> g;
        upsp_b200::read_plot3d_grid_file(name, g);
        report(g);
        if (argc > 3) upsp_b200::write_plot3d_grid_file(argv[3], g);
      }
      return 0;
    }
    auto g = upsp_b200::read_tri_grid(argv[1]);
    if (argc > 3 && std::string(argv[3]) == "intersect") {      // as psp_process loads it: TriModel_(file, intersect = true)
      const int n0 = g.n_nodes;
      const int overlap = upsp_b200::intersect_grid(g);
      std::printf("non_unique %d\nunique_overlapping %d\n", n0 - g.n_nodes, overlap);
    }
    std::vector<float> nrm;
    upsp_b200::calc_normals(g, nrm);
    std::vector<float> nrm_w;
    upsp_b200::node_normals_area_weighted(g, nrm_w);
    std::printf("n_nodes %d\nn_tris %d\nn_comps %d\nhas_comps %d\n", g.n_nodes, g.n_tris, g.number_of_components(),
                g.comps.empty() ? 0 : 1);
    if (argc > 2) {
      const std::string p = argv[2];
      auto dump = [&](const std::string& name, const void* d, size_t bytes) {
        FILE* f = std::fopen((p + name).c_str(), "wb");
        if (!f) throw std::runtime_error("cannot write " + p + name);
        std::fwrite(d, 1, bytes, f);
        std::fclose(f);
      };
      dump(".xyz", g.xyz.data(), g.xyz.size() * 4);
      dump(".nrm", nrm.data(), nrm.size() * 4);
      dump(".nrmw", nrm_w.data(), nrm_w.size() * 4);     // area-weighted (Node::get_normal): weights, target diameters
      dump(".tri", g.tris.data(), g.tris.size() * 4);
      dump(".comp", g.comps.data(), g.comps.size() * 4);
    }
  } catch (const std::exception& e) {
    std::cerr << "grid_probe: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
