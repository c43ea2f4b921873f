// grid_readers.hpp -- host-side readers of the model grid for the phase-0 products (SURVEY 8f
// rank 2).  Unstructured triangle grid (Cart3D-style Fortran-unformatted .tri), as
//   TriModel_<float>::load_grid   cpp/lib/TriModel.ipp:115-225   (records: {n_node, n_tri},
//       3*n_node float32 coordinates, 3*n_tri int32 1-based node ids, optional n_tri int32
//       component ids; every record framed by its int32 byte count; same error messages)
//   TriModel_<float>::calcNormals cpp/lib/TriModel.ipp:1428-1506 (node normal = normalised sum of the
//       unit normals of the triangles around the node, in ascending triangle order)
// The outputs are exactly what upsp_op_create_projection consumes: xyz, normals, triNodes.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
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

/* TriModel_::intersect_grid (cpp/lib/TriModel.ipp:941-1118), which psp_process runs on every .tri grid it loads
 * (TriModel_(file, intersect = true), psp_process.cpp:1384): nodes with bit-identical coordinates collapse onto the lowest
 * index of their group, triangles that collapse are dropped, nodes left without a triangle BETWEEN two removed nodes are
 * removed as well (the reference's own range logic, kept: an orphan before the first or after the last removed node stays),
 * and the remaining nodes are renumbered in order.  Returns the reference's "unique overlapping points" count; the number
 * of removed nodes is the change of n_nodes. */
inline int intersect_grid(TriGrid& g) {
  const int N = g.n_nodes;
  auto coord = [&](int n, int d) { return g.xyz[(size_t)n * 3 + d]; };
  std::vector<int> order((size_t)N);
  for (int n = 0; n < N; ++n) order[(size_t)n] = n;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    for (int d = 0; d < 3; ++d) {
      if (coord(a, d) < coord(b, d)) return true;
      if (coord(b, d) < coord(a, d)) return false;
    }
    return a < b;
  });
  std::vector<int> rep((size_t)N);
  int n_groups = 0;
  for (size_t i = 0; i < order.size();) {
    size_t j = i + 1;
    while (j < order.size() && coord(order[j], 0) == coord(order[i], 0) && coord(order[j], 1) == coord(order[i], 1) &&
           coord(order[j], 2) == coord(order[i], 2))
      ++j;
    for (size_t k = i; k < j; ++k) rep[(size_t)order[k]] = order[i];     // lowest index of the group (ties sorted by index)
    if (j - i > 1) ++n_groups;
    i = j;
  }
  std::vector<int> removed_sorted;                                         // the `second` of every duplicate pair that is kept
  for (int n = 0; n < N; ++n)
    if (rep[(size_t)n] != n) removed_sorted.push_back(n);
  if (removed_sorted.empty()) return 0;

  // move the triangles onto the representatives, drop the ones that collapse
  std::vector<int32_t> tris, comps;
  std::vector<int> n_faces((size_t)N, 0);
  for (int t = 0; t < g.n_tris; ++t) {
    const int32_t a = rep[(size_t)g.tris[(size_t)t * 3]], b = rep[(size_t)g.tris[(size_t)t * 3 + 1]], c = rep[(size_t)g.tris[(size_t)t * 3 + 2]];
    if (a == b || a == c || b == c) continue;
    tris.insert(tris.end(), {a, b, c});
    if (!g.comps.empty()) comps.push_back(g.comps[(size_t)t]);
    ++n_faces[(size_t)a];
    ++n_faces[(size_t)b];
    ++n_faces[(size_t)c];
  }
  // orphans between consecutive removed nodes (after the last one the reference's range ends at the NUMBER of duplicates)
  std::vector<char> removed((size_t)N, 0);
  for (int n : removed_sorted) removed[(size_t)n] = 1;
  int n_orphans = 0;
  for (size_t i = 0; i < removed_sorted.size(); ++i) {
    const int start = removed_sorted[i] + 1;
    const int end = i + 1 < removed_sorted.size() ? removed_sorted[i + 1] : (int)removed_sorted.size();
    for (int n = start; n < end; ++n)
      if (n_faces[(size_t)n] == 0 && !removed[(size_t)n]) removed[(size_t)n] = 1, ++n_orphans;
  }
  // renumber
  std::vector<int32_t> new_index((size_t)N, -1);
  std::vector<float> xyz;
  int next = 0;
  for (int n = 0; n < N; ++n)
    if (!removed[(size_t)n]) {
      new_index[(size_t)n] = next++;
      xyz.insert(xyz.end(), {coord(n, 0), coord(n, 1), coord(n, 2)});
    }
  for (int32_t& v : tris) v = new_index[(size_t)v];
  g.xyz.swap(xyz);
  g.tris.swap(tris);
  g.comps.swap(comps);
  g.n_nodes = next;
  g.n_tris = (int)(g.tris.size() / 3);
  return n_groups + n_orphans;
}

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

/* TriModel_::Node::get_normal (cpp/lib/TriModel.ipp:1570-1590): the AREA-weighted node normal, sum over the node's faces
 * (ascending face index) of upsp::normal(tri) * upsp::area(tri) (cpp/lib/models.ipp:137-182: cross product normalised with a
 * double norm; Heron's formula on the sorted float edge lengths), normalised.  This, not calc_normals, is what the reference's
 * camera weights (projection.ipp:990) and target diameters (psp_process.cpp:145) take for an unstructured model. */
inline void node_normals_area_weighted(const TriGrid& g, std::vector<float>& normals) {
  normals.assign((size_t)3 * g.n_nodes, 0.f);
  auto norm_d = [](float x, float y, float z) { return std::sqrt((double)x * x + (double)y * y + (double)z * z); };   // cv::norm
  for (int t = 0; t < g.n_tris; ++t) {
    const int32_t id[3] = {g.tris[3 * t], g.tris[3 * t + 1], g.tris[3 * t + 2]};
    const float* p0 = &g.xyz[3 * (size_t)id[0]];
    const float* p1 = &g.xyz[3 * (size_t)id[1]];
    const float* p2 = &g.xyz[3 * (size_t)id[2]];
    // upsp::normal: (p2 - p1) x (p0 - p1)
    const float ax = p2[0] - p1[0], ay = p2[1] - p1[1], az = p2[2] - p1[2];
    const float bx = p0[0] - p1[0], by = p0[1] - p1[1], bz = p0[2] - p1[2];
    float n[3] = {ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx};
    const double nn = norm_d(n[0], n[1], n[2]);
    if ((float)nn != 0.f)
      for (float& c : n) c = (float)((double)c / nn);              // Point3f / double
    // upsp::area: Heron, edges sorted a >= b >= c
    float a = (float)norm_d(p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]);
    float b = (float)norm_d(p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]);
    float c = (float)norm_d(p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]);
    if (b > a) std::swap(a, b);
    if (c > a) {
      const float tmp = a;
      a = c;
      c = b;
      b = tmp;
    } else if (c > b) {
      std::swap(b, c);
    }
    const float pos_neg = std::fabs(c - (a - b));
    const float area = (float)(0.25 * std::sqrt((a + (b + c)) * pos_neg * (c + (a - b)) * (a + (b - c))));
    for (int32_t node : id)
      for (int d = 0; d < 3; ++d) normals[3 * (size_t)node + d] += n[d] * area;
  }
  for (int node = 0; node < g.n_nodes; ++node) {
    float* v = &normals[3 * (size_t)node];
    const double m = norm_d(v[0], v[1], v[2]);
    if (m != 0.0)
      for (int d = 0; d < 3; ++d) v[d] = (float)((double)v[d] / m);
  }
}

// ---------------------------------------------------------------------------------------------
// Structured grids: unformatted plot3d (cpp/include/plot3d.h:30-110, read_plot3d_grid_file /
// write_plot3d_grid_file): Fortran records; single zone = {IDIM,JDIM,KDIM} then one x|y|z record;
// multi zone = {NZONES}, {3*NZONES dims}, one x|y|z[|iblank] record per zone; single or double
// precision, little or big endian, IBLANKs skipped.  The precision / endianness / IBLANK variant is
// recognised from the record lengths.  StructuredGrid<FP> mirrors upsp::StructuredGrid<FP>.
template <typename FP>
struct StructuredGrid {
  typedef FP data_type;
  std::vector<std::vector<unsigned>> grid_size;   // per zone {i, j, k}
  std::vector<FP> x, y, z;                        // all zones back to back, i fastest
  unsigned num_zones() const { return (unsigned)grid_size.size(); }
  unsigned zone_size(unsigned zn) const { return grid_size[zn][0] * grid_size[zn][1] * grid_size[zn][2]; }
  size_t size() const { return x.size(); }
};

namespace p3d_detail {
inline uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24); }
inline void swap_bytes(void* p, size_t elem, size_t n) {
  unsigned char* b = static_cast<unsigned char*>(p);
  for (size_t i = 0; i < n; ++i)
    for (size_t k = 0; k < elem / 2; ++k) std::swap(b[i * elem + k], b[i * elem + elem - 1 - k]);
}
}  // namespace p3d_detail

template <typename StructGrid>
void read_plot3d_grid_file(const std::string& filename, StructGrid& grid) {
  typedef typename StructGrid::data_type FP;
  std::ifstream ifs(filename, std::ios::in | std::ios::binary);
  if (!ifs) throw std::invalid_argument("Cannot open plot3d grid file '" + filename + "'");
  bool swap = false;
  auto rd_i32 = [&](int32_t* dst, size_t n) {
    ifs.read(reinterpret_cast<char*>(dst), (std::streamsize)(n * 4));
    if (!ifs) throw std::invalid_argument("Unable to read plot3d grid file '" + filename + "'");
    if (swap) p3d_detail::swap_bytes(dst, 4, n);
  };
  int32_t len = 0;
  rd_i32(&len, 1);
  if (len != 4 && len != 12) {                     // not a native-endian header record: try the other endianness
    len = (int32_t)p3d_detail::bswap32((uint32_t)len);
    swap = true;
    if (len != 4 && len != 12) throw std::invalid_argument("Unable to read plot3d grid file '" + filename + "': bad header record");
  }
  int32_t nz = 1, tail = 0;
  std::vector<int32_t> dims;
  if (len == 4) {                                   // multi-zone
    rd_i32(&nz, 1);
    rd_i32(&tail, 1);
    if (tail != len || nz <= 0) throw std::invalid_argument("Unable to read plot3d grid file, bad zone count");
    rd_i32(&len, 1);
    if (len != 12 * nz) throw std::invalid_argument("Unable to read plot3d grid file, inconsistent zone dimensions");
  }
  dims.resize((size_t)3 * nz);
  rd_i32(dims.data(), dims.size());
  rd_i32(&tail, 1);
  if (tail != len) throw std::invalid_argument("Unable to read plot3d grid file, inconsistent zone dimensions");
  grid.grid_size.clear();
  grid.x.clear(); grid.y.clear(); grid.z.clear();
  for (int zn = 0; zn < nz; ++zn) {
    if (dims[3 * zn] <= 0 || dims[3 * zn + 1] <= 0 || dims[3 * zn + 2] <= 0)
      throw std::invalid_argument("Unable to read plot3d grid file, non-positive zone dimension");
    grid.grid_size.push_back({(unsigned)dims[3 * zn], (unsigned)dims[3 * zn + 1], (unsigned)dims[3 * zn + 2]});
  }
  for (int zn = 0; zn < nz; ++zn) {
    const size_t npts = (size_t)grid.zone_size((unsigned)zn);
    uint32_t rlen = 0;
    ifs.read(reinterpret_cast<char*>(&rlen), 4);
    if (!ifs) throw std::invalid_argument("Unable to read plot3d grid file, missing zone data");
    if (swap) rlen = p3d_detail::bswap32(rlen);
    size_t elem = 0;
    bool iblank = false;
    if (rlen == 12 * npts) elem = 4;
    else if (rlen == 24 * npts) elem = 8;
    else if (rlen == 16 * npts) { elem = 4; iblank = true; }
    else if (rlen == 28 * npts) { elem = 8; iblank = true; }
    else throw std::invalid_argument("Unable to read plot3d grid file, zone record length matches no precision");
    std::vector<unsigned char> buf(3 * npts * elem);
    ifs.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)buf.size());
    if (!ifs) throw std::invalid_argument("Unable to read plot3d grid file, truncated zone data");
    if (swap) p3d_detail::swap_bytes(buf.data(), elem, 3 * npts);
    if (iblank) ifs.seekg((std::streamoff)(4 * npts), std::ios::cur);
    uint32_t rtail = 0;
    ifs.read(reinterpret_cast<char*>(&rtail), 4);
    if (swap) rtail = p3d_detail::bswap32(rtail);
    if (!ifs || rtail != rlen) throw std::invalid_argument("Unable to read plot3d grid file, inconsistent zone record");
    std::vector<FP>* out[3] = {&grid.x, &grid.y, &grid.z};
    for (int c = 0; c < 3; ++c)
      for (size_t i = 0; i < npts; ++i) {
        if (elem == 4) {
          float v;
          std::memcpy(&v, &buf[(c * npts + i) * 4], 4);
          out[c]->push_back((FP)v);
        } else {
          double v;
          std::memcpy(&v, &buf[(c * npts + i) * 8], 8);
          out[c]->push_back((FP)v);
        }
      }
  }
}

/* machine-endian, precision of StructGrid::data_type, single-zone layout when there is one zone */
template <typename StructGrid>
void write_plot3d_grid_file(const std::string& filename, const StructGrid& grid) {
  typedef typename StructGrid::data_type FP;
  std::ofstream ofs(filename, std::ios::out | std::ios::binary);
  if (!ofs) throw std::invalid_argument("Cannot open plot3d grid file '" + filename + "' for writing");
  auto rec = [&](const void* p, size_t bytes) {
    const int32_t n = (int32_t)bytes;
    ofs.write(reinterpret_cast<const char*>(&n), 4);
    ofs.write(reinterpret_cast<const char*>(p), (std::streamsize)bytes);
    ofs.write(reinterpret_cast<const char*>(&n), 4);
  };
  const int32_t nz = (int32_t)grid.num_zones();
  if (nz != 1) rec(&nz, 4);
  std::vector<int32_t> dims;
  for (const auto& g : grid.grid_size) dims.insert(dims.end(), {(int32_t)g[0], (int32_t)g[1], (int32_t)g[2]});
  rec(dims.data(), dims.size() * 4);
  size_t off = 0;
  for (int32_t zn = 0; zn < nz; ++zn) {
    const size_t npts = grid.zone_size((unsigned)zn);
    std::vector<FP> buf;
    buf.insert(buf.end(), grid.x.begin() + (std::ptrdiff_t)off, grid.x.begin() + (std::ptrdiff_t)(off + npts));
    buf.insert(buf.end(), grid.y.begin() + (std::ptrdiff_t)off, grid.y.begin() + (std::ptrdiff_t)(off + npts));
    buf.insert(buf.end(), grid.z.begin() + (std::ptrdiff_t)off, grid.z.begin() + (std::ptrdiff_t)(off + npts));
    rec(buf.data(), buf.size() * sizeof(FP));
    off += npts;
  }
}

}  // namespace upsp_b200
