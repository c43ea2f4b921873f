// p3d_model.hpp -- the structured (plot3d) surface model of psp_process, host side.
// What the frame chain needs from upsp::P3DModel_<float> (cpp/lib/P3DModel.ipp), on plain arrays:
//   identify_overlap     :893-1127   zone-edge nodes within `tol` of each other (other zone, or the same zone when it
//                                    wraps) -> overlap_pts: node -> sorted list of the nodes it overlaps
//   get_low_nidx / is_superceded / is_overlapping   :692-703, 1346-1354
//   adjust_solution      :144-157    sol[alt] = sol[curr] for curr < alt, map order
//   overlap_src_index                the same loop run once on 0..N-1: out[n] = in[src_index[n]] is what
//                                    upsp_gpu_set_overlap_remap takes (row a8 of DESIGN.md)
//   extract_tris         :234-317    the two triangles of every quad, node indices as fed to the ray caster
//   calc_normals         :1357-1654  node normal = normalised sum of the unit normals of the <= 4 faces around the node
//                                    and around every node that overlaps it
// The reference finds neighbours with a kd-tree range query (cpp/raycast/pspKdtree.c:225-256: double
// coordinates, squared distance <= range^2); here the edge nodes are sorted along the widest axis and swept, with the
// same predicate.  Compile with -ffp-contract=off.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <set>
#include <stdexcept>
#include <vector>

#include "grid_readers.hpp"

namespace upsp_b200 {

struct GridIndex {
  int zone = 0, j = 0, k = 0;
};

class P3DModel {
 public:
  typedef int node_idx;

  template <typename FPin>
  explicit P3DModel(const StructuredGrid<FPin>& grid, float tol = 0.f) { load_grid(grid, tol); }
  explicit P3DModel(const std::string& filename, float tol = 0.f) {
    StructuredGrid<float> g;
    read_plot3d_grid_file(filename, g);
    load_grid(g, tol);
  }

  template <typename FPin>
  void load_grid(const StructuredGrid<FPin>& grid, float tol) {
    overlap_tol_ = tol;
    x_.assign(grid.x.begin(), grid.x.end());
    y_.assign(grid.y.begin(), grid.y.end());
    z_.assign(grid.z.begin(), grid.z.end());
    sz_.clear();
    start_.clear();
    int base = 0;
    for (unsigned zn = 0; zn < grid.num_zones(); ++zn) {
      const auto& s = grid.grid_size[zn];
      if (s[2] != 1) throw std::invalid_argument("P3DModel: zone " + std::to_string(zn) + " is not a j-k surface (l != 1)");
      sz_.push_back({(int)s[0], (int)s[1]});
      start_.push_back(base);
      base += (int)(s[0] * s[1]);
    }
    if ((size_t)base != x_.size()) throw std::invalid_argument("P3DModel: zone sizes inconsistent with the point count");
    const auto res = identify_overlap(overlap_tol_);
    n_vert_ = (int)x_.size() - res.first + res.second;
    n_face_ = 0;
    for (const auto& s : sz_) n_face_ += (s.j - 1) * (s.k - 1);
    calc_normals();
  }

  int size() const { return (int)x_.size(); }
  int num_zones() const { return (int)sz_.size(); }
  int zone_size(int zn) const { return sz_[zn].j * sz_[zn].k; }
  int zone_size(int zn, int dir) const { return dir == 0 ? sz_[zn].j : sz_[zn].k; }
  int zone_start_idx(int zn) const { return start_[zn]; }
  int number_of_faces() const { return n_face_; }
  int number_of_vertices() const { return n_vert_; }   // nodes that are not superceded
  const std::vector<float>& get_x() const { return x_; }
  const std::vector<float>& get_y() const { return y_; }
  const std::vector<float>& get_z() const { return z_; }
  const std::vector<float>& get_n() const { return normals_; }   // [N][3]
  const std::map<node_idx, std::vector<node_idx>>& overlap_pts() const { return overlap_pts_; }

  GridIndex nidx2_gidx(node_idx nidx) const {
    int zn = (int)(std::upper_bound(start_.begin(), start_.end(), nidx) - start_.begin()) - 1;
    const int idx = nidx - start_[zn];
    return {zn, idx % sz_[zn].j, idx / sz_[zn].j};
  }
  node_idx gidx2_nidx(const GridIndex& g) const { return start_[g.zone] + g.k * sz_[g.zone].j + g.j; }

  bool is_overlapping(node_idx nidx) const { return overlap_pts_.count(nidx) != 0; }
  node_idx get_low_nidx(node_idx nidx) const {
    auto it = overlap_pts_.find(nidx);
    if (it != overlap_pts_.end() && it->second[0] < nidx) return it->second[0];
    return nidx;
  }
  bool is_superceded(node_idx nidx) const { return get_low_nidx(nidx) != nidx; }

  template <typename T>
  void adjust_solution(std::vector<T>& sol) const {
    for (const auto& kv : overlap_pts_)
      for (node_idx alt : kv.second)
        if (kv.first < alt) sol[alt] = sol[kv.first];
  }
  /* static form of adjust_solution: adjusted[n] = original[src_index[n]] */
  std::vector<int32_t> overlap_src_index() const {
    std::vector<int32_t> idx(x_.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int32_t)i;
    adjust_solution(idx);
    return idx;
  }

  /* triangles handed to the ray caster: per quad (i0,i1,i2), (i2,i3,i0) */
  void extract_tris(std::vector<float>& tris, std::vector<int>& tri_nodes) const {
    tris.clear();
    tri_nodes.clear();
    for (int zn = 0; zn < num_zones(); ++zn) {
      const int J = sz_[zn].j, K = sz_[zn].k, base = start_[zn];
      for (int q = 0; q < (J - 1) * (K - 1); ++q) {
        const int klo = q / (J - 1), jlo = q % (J - 1);
        const int i0 = base + klo * J + jlo, i1 = i0 + 1, i2 = i1 + J, i3 = i0 + J;
        for (int n : {i0, i1, i2, i2, i3, i0}) {
          tri_nodes.push_back(n);
          tris.push_back(x_[n]);
          tris.push_back(y_[n]);
          tris.push_back(z_[n]);
        }
      }
    }
  }

 private:
  struct ZoneSize {
    int j, k;
  };

  bool is_edge(int zn, int idx) const {
    const int r = idx / sz_[zn].j, c = idx % sz_[zn].j;
    return r == 0 || r == sz_[zn].k - 1 || c == 0 || c == sz_[zn].j - 1;
  }

  /* returns (#nodes that take part in any overlap, #overlap groups as the reference counts them) */
  std::pair<int, int> identify_overlap(float tol) {
    const float min_tol = 1e-12f;
    tol = tol > min_tol ? tol : min_tol;
    const double range = (double)tol, range_sq = range * range;
    overlap_pts_.clear();

    std::vector<int> edge;   // ascending nidx
    for (int zn = 0; zn < num_zones(); ++zn)
      for (int idx = 0; idx < zone_size(zn); ++idx)
        if (is_edge(zn, idx)) edge.push_back(start_[zn] + idx);
    if (edge.empty()) return {0, 0};

    // sweep axis = widest extent of the edge nodes
    const std::vector<float>* ax[3] = {&x_, &y_, &z_};
    int a = 0;
    double best = -1;
    for (int d = 0; d < 3; ++d) {
      float lo = (*ax[d])[edge[0]], hi = lo;
      for (int n : edge) lo = std::min(lo, (*ax[d])[n]), hi = std::max(hi, (*ax[d])[n]);
      if ((double)hi - lo > best) best = (double)hi - lo, a = d;
    }
    const std::vector<float>& key = *ax[a];
    std::vector<int> order(edge);
    std::sort(order.begin(), order.end(), [&](int p, int q) { return key[p] < key[q] || (key[p] == key[q] && p < q); });

    // others[n] for every edge node, found from both sides of the symmetric predicate
    std::map<int, std::set<int>> others;
    for (size_t p = 0; p < order.size(); ++p) {
      const int n = order[p];
      const GridIndex g = nidx2_gidx(n);
      for (size_t q = p + 1; q < order.size(); ++q) {
        const int m = order[q];
        const double dk = (double)key[m] - (double)key[n];
        if (dk > range) break;
        const double dx = (double)x_[m] - x_[n], dy = (double)y_[m] - y_[n], dz = (double)z_[m] - z_[n];
        if (dx * dx + dy * dy + dz * dz > range_sq) continue;
        const GridIndex o = nidx2_gidx(m);
        if (o.zone == g.zone) {   // only a wrapped zone overlaps itself
          bool wrapped = false;
          const ZoneSize d = sz_[g.zone];
          if (g.j == o.j) wrapped = (g.k == 0 && o.k == d.k - 1) || (g.k == d.k - 1 && o.k == 0);
          else if (g.k == o.k) wrapped = (g.j == 0 && o.j == d.j - 1) || (g.j == d.j - 1 && o.j == 0);
          if (!wrapped) continue;
        }
        others[n].insert(m);
        others[m].insert(n);
      }
    }

    // the reference's bookkeeping, edge nodes in ascending order
    std::set<int> oset, seen, uniq;
    for (int n : edge) {
      auto it = others.find(n);
      if (it == others.end()) continue;
      if (!seen.count(n) && !uniq.count(n)) {
        uniq.insert(n);
        seen.insert(n);
      }
      for (int m : it->second) {
        seen.insert(m);
        oset.insert(n);
        oset.insert(m);
      }
      overlap_pts_[n] = std::vector<int>(it->second.begin(), it->second.end());   // sorted, unique, never n itself
    }
    return {(int)oset.size(), (int)uniq.size()};
  }

  static void cross_uv(const float* p0, const float* p1, const float* p2, float n[3]) {
    const float ux = p2[0] - p1[0], uy = p2[1] - p1[1], uz = p2[2] - p1[2];
    const float vx = p0[0] - p1[0], vy = p0[1] - p1[1], vz = p0[2] - p1[2];
    n[0] = uy * vz - vy * uz;
    n[1] = vx * uz - ux * vz;
    n[2] = ux * vy - vx * uy;
  }
  static float norm3(const float n[3]) {   // cv::norm(Point3f): double accumulate, narrowed by the caller
    return (float)std::sqrt((double)n[0] * n[0] + (double)n[1] * n[1] + (double)n[2] * n[2]);
  }

  void add_faces_around(node_idx nidx, float acc[3]) const {
    const GridIndex g = nidx2_gidx(nidx);
    const int J = sz_[g.zone].j, K = sz_[g.zone].k, base = start_[g.zone];
    const int row = g.k, col = g.j;
    const int cm = base + row * J + std::max(0, col - 1), cp = base + row * J + std::min(J - 1, col + 1);
    const int rm = base + std::max(0, row - 1) * J + col, rp = base + std::min(K - 1, row + 1) * J + col;
    // LL, LR, UR, UL: (v0, v1 = node, v2); the fourth corner only enters the (unused) area
    const int quads[4][2] = {{rm, cm}, {cp, rm}, {rp, cp}, {cm, rp}};
    for (const auto& q : quads) {
      if (q[0] == nidx || q[1] == nidx) continue;   // no face on that side
      float p0[3] = {x_[q[0]], y_[q[0]], z_[q[0]]}, p1[3] = {x_[nidx], y_[nidx], z_[nidx]}, p2[3] = {x_[q[1]], y_[q[1]], z_[q[1]]};
      float n[3];
      cross_uv(p0, p1, p2, n);
      const float mag = norm3(n);
      for (int d = 0; d < 3; ++d) acc[d] += (mag == 0.f) ? n[d] : n[d] / mag;
    }
  }

  void calc_normals() {
    const int N = size();
    normals_.assign((size_t)N * 3, 0.f);
    for (int nidx = 0; nidx < N; ++nidx) {
      float acc[3] = {0.f, 0.f, 0.f};
      add_faces_around(nidx, acc);
      auto it = overlap_pts_.find(nidx);
      if (it != overlap_pts_.end())
        for (node_idx o : it->second) add_faces_around(o, acc);
      const float mag = norm3(acc);
      for (int d = 0; d < 3; ++d) normals_[(size_t)nidx * 3 + d] = (mag == 0.f) ? acc[d] : acc[d] / mag;
    }
  }

  float overlap_tol_ = 0.f;
  std::vector<float> x_, y_, z_, normals_;
  std::vector<ZoneSize> sz_;
  std::vector<int> start_;
  std::map<node_idx, std::vector<node_idx>> overlap_pts_;
  int n_vert_ = 0, n_face_ = 0;
};

}  // namespace upsp_b200
