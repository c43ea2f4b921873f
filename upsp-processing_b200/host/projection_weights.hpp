// projection_weights.hpp -- multi-camera blending weights of the projection matrices, host side
// (the values that the frame chain's camera blend, a7, then simply sums).
// Mirrors, on plain CSR arrays,
//   adjust_projection_for_weights   cpp/lib/projection.ipp:912-1078  (nodes seen by several cameras:
//       every camera's row is scaled by its weight; the cameras of a node are visited in the order a
//       std::priority_queue with the reference's comparator pops them -- the same container, the same
//       comparator and the same push sequence are used here, so that order is reproduced)
//   BestView / AverageViews         cpp/lib/projection.ipp:226-268
//   angle_between                   cpp/utils/cv_extras.ipp:67-73 (float dot product, double norms, acos)
//   identify_skipped_nodes          cpp/lib/projection.ipp:857-880
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <queue>
#include <vector>

namespace upsp_b200 {

struct CsrMatrix {                 // Eigen::SparseMatrix<float, RowMajor>, compressed
  std::vector<int32_t> rowptr, col;
  std::vector<float> val;
  int rows() const { return (int)rowptr.size() - 1; }
};

enum class OverlapType { BestView, AverageViews };   // deck @options overlap (upsp_inputs.h)

inline std::vector<float> best_view(const std::vector<float>& angles) {
  std::vector<float> w(angles.size(), 0.f);
  if (angles.empty()) return w;
  size_t mi = 0;
  for (size_t i = 1; i < angles.size(); ++i)
    if (angles[i] > angles[mi]) mi = i;
  w[mi] = 1.f;
  return w;
}
inline std::vector<float> average_views(const std::vector<float>& angles) {
  std::vector<float> w(angles.size(), 0.f);
  float sum = 0.0f;
  for (float a : angles) sum += a;
  for (size_t i = 0; i < angles.size(); ++i) w[i] = angles[i] / sum;
  return w;
}

inline double angle_between(const float v1[3], const float v2[3]) {
  auto norm = [](const float v[3]) { return std::sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]); };
  const double ang = (double)(v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]) / norm(v1) / norm(v2);
  return std::acos(ang);
}

/* xyz, normals: [n_nodes][3] (Node::get_position / get_normal); centers: CameraCal::get_cam_center per camera */
inline void adjust_projection_for_weights(const float* xyz, const float* normals, const std::vector<std::array<double, 3>>& centers,
                                          std::vector<CsrMatrix>& projs, OverlapType overlap) {
  struct ProjWeights {
    int row = 0, cam = 0;
    CsrMatrix* proj = nullptr;
    float center[3];
    void first() {
      for (row = 0; row < proj->rows(); ++row)
        if (proj->rowptr[(size_t)row + 1] > proj->rowptr[(size_t)row]) return;
      row = -1;
    }
    void next() {
      if (row < 0) return;
      for (int k = row + 1; k < proj->rows(); ++k)
        if (proj->rowptr[(size_t)k + 1] > proj->rowptr[(size_t)k]) {
          row = k;
          return;
        }
      row = -1;
    }
    bool greater(const ProjWeights& o) const {
      if (o.row < 0) return false;
      if (row < 0) return true;
      return row > o.row;
    }
    bool equal(const ProjWeights& o) const { return o.row < 0 ? row < 0 : row == o.row; }
  };
  struct Cmp {
    bool operator()(const ProjWeights* a, const ProjWeights* b) const { return a->greater(*b); }
  };
  std::vector<ProjWeights> pws(projs.size());
  std::priority_queue<ProjWeights*, std::vector<ProjWeights*>, Cmp> qu;
  for (size_t c = 0; c < projs.size(); ++c) {
    pws[c].cam = (int)c;
    pws[c].proj = &projs[c];
    for (int i = 0; i < 3; ++i) pws[c].center[i] = (float)centers[c][(size_t)i];
    pws[c].first();
    if (pws[c].row >= 0) qu.push(&pws[c]);
  }
  auto angle_of = [&](const ProjWeights* p) {
    const float* pos = xyz + 3 * (size_t)p->row;
    const float dir[3] = {pos[0] - p->center[0], pos[1] - p->center[1], pos[2] - p->center[2]};
    return (float)angle_between(dir, normals + 3 * (size_t)p->row);
  };
  while (!qu.empty()) {
    ProjWeights* pw = qu.top();
    qu.pop();
    if (qu.empty()) break;
    if (qu.top()->equal(*pw)) {
      std::vector<ProjWeights*> grp(1, pw);
      std::vector<float> angs(1, angle_of(pw));
      while (qu.top()->equal(*pw)) {
        grp.push_back(qu.top());
        qu.pop();
        angs.push_back(angle_of(grp.back()));
        if (qu.empty()) break;
      }
      const std::vector<float> w = overlap == OverlapType::BestView ? best_view(angs) : average_views(angs);
      for (size_t i = 0; i < grp.size(); ++i) {
        CsrMatrix& m = *grp[i]->proj;
        for (int32_t k = m.rowptr[(size_t)grp[i]->row]; k < m.rowptr[(size_t)grp[i]->row + 1]; ++k) m.val[(size_t)k] *= w[i];
        grp[i]->next();
        if (grp[i]->row >= 0) qu.push(grp[i]);
      }
    } else {
      pw->next();
      if (pw->row >= 0) qu.push(pw);
    }
  }
}

inline void identify_skipped_nodes(const std::vector<CsrMatrix>& projs, std::vector<unsigned>& skipped) {
  skipped.clear();
  if (projs.empty()) return;
  for (int i = 0; i < projs[0].rows(); ++i) {
    bool found = false;
    for (const auto& m : projs)
      if (m.rowptr[(size_t)i + 1] > m.rowptr[(size_t)i]) { found = true; break; }
    if (!found) skipped.push_back((unsigned)i);
  }
}

}  // namespace upsp_b200
