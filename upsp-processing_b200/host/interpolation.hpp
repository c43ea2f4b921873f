// interpolation.hpp -- steady-state Cp / model temperature from a structured grid onto the nodes of an unstructured
// model, host side.  Mirrors upsp::interpolate (cpp/lib/interpolation.ipp:17-67) as phase 2 calls it
// (cpp/exec/psp_process.cpp:2338-2345, 2371-2378: k = 10, p = 2.0, steady grid loaded with tol 1e-3): for every output
// node the k nearest valid (not superceded) nodes of the input model, inverse-distance weights 1 / dist^p in float, an
// exact hit takes that node's value.  The reference finds the neighbours with an octree; here a uniform cell grid over
// the input nodes is searched in growing shells, which returns the same k nodes (ties in distance: lowest index).  The
// order in which the k terms are summed (here: nearest first) is the octree's in the reference and is not defined by its
// interface: parity unpinned in the last bits.  Compile with -ffp-contract=off.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "p3d_model.hpp"

namespace upsp_b200 {

class NodeCellGrid {
 public:
  /* xyz [N][3]; valid[n] != 0 marks the nodes that may be returned */
  NodeCellGrid(const float* xyz, int n, const uint8_t* valid) : xyz_(xyz) {
    for (int d = 0; d < 3; ++d) lo_[d] = 1e300, hi_[d] = -1e300;
    int nv = 0;
    for (int i = 0; i < n; ++i)
      if (valid[i]) {
        ++nv;
        for (int d = 0; d < 3; ++d) lo_[d] = std::min(lo_[d], (double)xyz[3 * i + d]), hi_[d] = std::max(hi_[d], (double)xyz[3 * i + d]);
      }
    if (nv == 0) throw std::invalid_argument("NodeCellGrid: no valid nodes");
    const double ext = std::max({hi_[0] - lo_[0], hi_[1] - lo_[1], hi_[2] - lo_[2], 1e-30});
    const int per_axis = std::max(1, std::min(256, (int)std::ceil(2.0 * std::cbrt((double)nv))));
    cell_ = ext / per_axis;
    for (int d = 0; d < 3; ++d) dim_[d] = std::max(1, std::min(per_axis, (int)std::floor((hi_[d] - lo_[d]) / cell_) + 1));
    start_.assign((size_t)dim_[0] * dim_[1] * dim_[2] + 1, 0);
    std::vector<int> cell_of((size_t)n, -1);
    for (int i = 0; i < n; ++i)
      if (valid[i]) {
        int c[3];
        cell_coords(xyz + 3 * i, c);
        cell_of[(size_t)i] = (c[2] * dim_[1] + c[1]) * dim_[0] + c[0];
        ++start_[(size_t)cell_of[(size_t)i] + 1];
      }
    for (size_t c = 1; c < start_.size(); ++c) start_[c] += start_[c - 1];
    items_.resize((size_t)nv);
    std::vector<int> fill(start_.begin(), start_.end() - 1);
    for (int i = 0; i < n; ++i)
      if (cell_of[(size_t)i] >= 0) items_[(size_t)fill[(size_t)cell_of[(size_t)i]]++] = i;    // ascending index within a cell
  }

  /* the k nearest valid nodes of p, nearest first (ties: lowest index); fewer if the grid holds fewer */
  void nearest_k(const float p[3], unsigned k, std::vector<std::pair<double, int>>& best) const {
    best.clear();
    int c0[3];
    cell_coords(p, c0);
    const int max_r = std::max({dim_[0], dim_[1], dim_[2]});
    auto worse = [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a < b; };   // max-heap on (dist2, idx)
    for (int r = 0; r <= max_r; ++r) {
      if (best.size() == k) {
        const double reach = (double)(r - 1) * cell_;       // every cell outside shell r-1 is at least this far away
        if (reach > 0 && best.front().first < reach * reach) break;
      }
      for (int z = c0[2] - r; z <= c0[2] + r; ++z) {
        if (z < 0 || z >= dim_[2]) continue;
        for (int y = c0[1] - r; y <= c0[1] + r; ++y) {
          if (y < 0 || y >= dim_[1]) continue;
          const bool face = (z == c0[2] - r || z == c0[2] + r || y == c0[1] - r || y == c0[1] + r);
          const int step = face ? 1 : std::max(1, 2 * r);    // interior rows of the shell: only the two end cells
          for (int x = c0[0] - r; x <= c0[0] + r; x += step) {
            if (x < 0 || x >= dim_[0]) continue;
            const size_t c = ((size_t)z * dim_[1] + y) * dim_[0] + x;
            for (int it = start_[c]; it < start_[c + 1]; ++it) {
              const int n = items_[(size_t)it];
              const double dx = (double)xyz_[3 * n] - p[0], dy = (double)xyz_[3 * n + 1] - p[1], dz = (double)xyz_[3 * n + 2] - p[2];
              const std::pair<double, int> cand(dx * dx + dy * dy + dz * dz, n);
              if (best.size() < k) {
                best.push_back(cand);
                std::push_heap(best.begin(), best.end(), worse);
              } else if (cand < best.front()) {
                std::pop_heap(best.begin(), best.end(), worse);
                best.back() = cand;
                std::push_heap(best.begin(), best.end(), worse);
              }
            }
          }
        }
      }
    }
    std::sort(best.begin(), best.end());
  }

 private:
  void cell_coords(const float p[3], int c[3]) const {
    for (int d = 0; d < 3; ++d) {
      const double f = std::floor(((double)p[d] - lo_[d]) / cell_);
      c[d] = f < 0 ? 0 : (f >= dim_[d] ? dim_[d] - 1 : (int)f);     // queries outside the box start from its nearest cell
    }
  }
  const float* xyz_;
  double lo_[3], hi_[3], cell_ = 1;
  int dim_[3] = {1, 1, 1};
  std::vector<int> start_, items_;
};

/* data: one value per node of in_model; out_xyz [n_out][3] */
inline std::vector<float> interpolate(const P3DModel& in_model, const std::vector<float>& data, const float* out_xyz, int n_out,
                                      unsigned k = 10, float p = 2.0f) {
  const int N = in_model.size();
  if ((int)data.size() != N) throw std::invalid_argument("interpolate: data inconsistent with the input model");
  std::vector<float> xyz((size_t)N * 3);
  std::vector<uint8_t> valid((size_t)N);
  for (int n = 0; n < N; ++n) {
    xyz[(size_t)n * 3] = in_model.get_x()[(size_t)n];
    xyz[(size_t)n * 3 + 1] = in_model.get_y()[(size_t)n];
    xyz[(size_t)n * 3 + 2] = in_model.get_z()[(size_t)n];
    valid[(size_t)n] = in_model.is_superceded(n) ? 0 : 1;    // cnode_begin .. cnode_end
  }
  const NodeCellGrid grid(xyz.data(), N, valid.data());
  std::vector<float> out((size_t)n_out, 0.0f);
  std::vector<std::pair<double, int>> nodes;
  for (int i = 0; i < n_out; ++i) {
    const float* pt = out_xyz + 3 * (size_t)i;
    grid.nearest_k(pt, k, nodes);
    float total_weight = 0.0f;
    for (const auto& nd : nodes) {
      const float* pt2 = xyz.data() + 3 * (size_t)nd.second;
      const float d[3] = {pt2[0] - pt[0], pt2[1] - pt[1], pt2[2] - pt[2]};
      const float dist = (float)std::sqrt((double)d[0] * d[0] + (double)d[1] * d[1] + (double)d[2] * d[2]);   // cv::norm
      if (dist == 0.0f) {
        total_weight = 1.0f;
        out[(size_t)i] = data[(size_t)nd.second];
        break;
      }
      const float weight = (float)(1.0 / std::pow((double)dist, (double)p));
      out[(size_t)i] += data[(size_t)nd.second] * weight;
      total_weight += weight;
    }
    out[(size_t)i] /= total_weight;
  }
  return out;
}

}  // namespace upsp_b200
