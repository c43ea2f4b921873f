// grid_readers.hpp -- host-side readers of the model grid for the phase-0 products (SURVEY 8f
// rank 2).  Unstructured triangle grid (Cart3D-style Fortran-unformatted .tri), as
//   TriModel_<float>::load_grid   cpp/lib/TriModel.ipp:115-225   (records: {n_node, n_tri},
//       3*n_node float32 coordinates, 3*n_tri int32 1-based node ids, optional n_tri int32
//       component ids; every record framed by its int32 byte count; same error messages)
//   TriModel_<float>::calcNormals cpp/lib/TriModel.ipp:1428-1506 (node normal = normalised sum of the
//       unit normals of the triangles around the node, in ascending triangle order)
// The outputs are exactly what upsp_op_create_projection consumes: xyz, normals, triNodes.
#pragma once
#include <cmath>
#include <cstdint>
#include <fstream>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace upsp_b200 {

struct TriGrid {
  int n_nodes = 0, n_tris = 0;
  std::vector<float> xyz;        // [n_nodes][3]
  std::vector<int32_t> tris;     // [n_tris][3], 0-based node indices (the reference's triNodes)
  std::vector<int32_t> comps;    // [n_tris] component ids, empty if the file has none
  int number_of_components() const {
    if (comps.empty()) return 1;
    return (int)std::set<int32_t>(comps.begin(), comps.end()).size();
  }
};

inline TriGrid read_tri_grid(const std::string& model_file) {
  std::ifstream ifs(model_file, std::ios::in | std::ios::binary);
  if (!ifs) throw std::invalid_argument("Cannot open tri grid file '" + model_file + "'");
  auto rd32 = [&]() {
    int32_t v = 0;
    ifs.read(reinterpret_cast<char*>(&v), sizeof v);
    return v;
  };
  TriGrid g;
  int32_t sz = rd32();
  if (!ifs || sz != (int32_t)(2 * sizeof(int32_t))) throw std::invalid_argument("Unable to read tri grid file '" + model_file + "'");
  g.n_nodes = rd32();
  g.n_tris = rd32();
  if (rd32() != sz || g.n_nodes < 0 || g.n_tris < 0) throw std::invalid_argument("Unable to read tri grid file");
  sz = rd32();
  if ((int64_t)sz != (int64_t)sizeof(float) * 3 * g.n_nodes)
    throw std::invalid_argument("Unable to read tri grid file, inconsistent number of nodes");
  g.xyz.resize((size_t)3 * g.n_nodes);
  ifs.read(reinterpret_cast<char*>(g.xyz.data()), (std::streamsize)(g.xyz.size() * sizeof(float)));
  if (rd32() != sz || !ifs) throw std::invalid_argument("Unable to read tri grid file, inconsistent number of nodes");
  sz = rd32();
  if ((int64_t)sz != (int64_t)sizeof(int32_t) * 3 * g.n_tris)
    throw std::invalid_argument("Unable to read tri grid file, inconsistent number of faces");
  g.tris.resize((size_t)3 * g.n_tris);
  ifs.read(reinterpret_cast<char*>(g.tris.data()), (std::streamsize)(g.tris.size() * sizeof(int32_t)));
  if (rd32() != sz || !ifs) throw std::invalid_argument("Unable to read tri grid file, inconsistent number of faces");
  for (auto& n : g.tris) {
    n -= 1;                                        // 1-based in the file (TriModel.ipp:190-192)
    if (n < 0 || n >= g.n_nodes) throw std::invalid_argument("Unable to read tri grid file, face references a missing node");
  }
  sz = rd32();                                     // optional component record
  if (ifs) {
    if ((int64_t)sz != (int64_t)sizeof(int32_t) * g.n_tris)
      throw std::invalid_argument("Unable to read tri grid file, inconsistent number of face components");
    g.comps.resize((size_t)g.n_tris);
    ifs.read(reinterpret_cast<char*>(g.comps.data()), (std::streamsize)(g.comps.size() * sizeof(int32_t)));
    if (rd32() != sz || !ifs) throw std::invalid_argument("Unable to read tri grid file, inconsistent number of face components");
  }
  return g;
}

/* normals [n_nodes][3]; nodes without a (non-degenerate) triangle get (0,0,0) */
inline void calc_normals(const TriGrid& g, std::vector<float>& normals) {
  normals.assign((size_t)3 * g.n_nodes, 0.f);
  auto norm3 = [](float x, float y, float z) {     // cv::norm(Point3f): double accumulation, narrowed to float
    return (float)std::sqrt((double)x * x + (double)y * y + (double)z * z);
  };
  for (int t = 0; t < g.n_tris; ++t) {             // ascending triangle order == the order of n2t_'s std::set
    const int32_t n0 = g.tris[3 * t], n1 = g.tris[3 * t + 1], n2 = g.tris[3 * t + 2];
    const float* p0 = &g.xyz[3 * (size_t)n0];
    const float* p1 = &g.xyz[3 * (size_t)n1];
    const float* p2 = &g.xyz[3 * (size_t)n2];
    const float ux = p2[0] - p1[0], uy = p2[1] - p1[1], uz = p2[2] - p1[2];
    const float vx = p0[0] - p1[0], vy = p0[1] - p1[1], vz = p0[2] - p1[2];
    float nx = uy * vz - vy * uz, ny = vx * uz - ux * vz, nz = ux * vy - vx * uy;
    const float m = norm3(nx, ny, nz);
    if (m != 0.f) { nx /= m; ny /= m; nz /= m; }
    for (int32_t n : {n0, n1, n2}) {
      normals[3 * (size_t)n] += nx;
      normals[3 * (size_t)n + 1] += ny;
      normals[3 * (size_t)n + 2] += nz;
    }
  }
  for (int n = 0; n < g.n_nodes; ++n) {
    float* v = &normals[3 * (size_t)n];
    const float m = norm3(v[0], v[1], v[2]);
    if (m != 0.f) { v[0] /= m; v[1] /= m; v[2] /= m; }
  }
}

}  // namespace upsp_b200
