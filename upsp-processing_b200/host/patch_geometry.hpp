// patch_geometry.hpp -- phase-0 fiducial patch geometry on the host: from the
// projected targets (u, v, diameter) of one camera to the boundary / interior pixel lists that
// upsp_gpu_set_patches consumes.  Mirrors, with the same names and argument meaning,
//   cluster_points          cpp/lib/patches.ipp:240-275
//   get_target_boundary     cpp/lib/patches.ipp:279-326
//   get_cluster_boundary    cpp/lib/patches.ipp:330-487
//   PatchClusters ctor      cpp/lib/patches.ipp:15-54   (pixels outside the frame are dropped)
//   threshold_bounds        cpp/lib/patches.ipp:59-94   (boundary pixels whose (2*offset+1)^2
//                           neighbourhood of the reference frame dips below `thresh` are dropped)
// as called from InitializeImagePatches (cpp/exec/psp_process.cpp:2125-2163).  Pure integer / float
// logic, no OpenCV / Eigen / Boost.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <list>
#include <queue>
#include <vector>

namespace upsp_b200 {

struct Target {            // upsp::Target_<float>: image position and diameter in pixels
  float u = 0.f, v = 0.f, diameter = 0.f;
};
struct Point2i {
  int x = 0, y = 0;
};

/* group targets whose patches would touch: breadth-first over "closer than bound_pts + mean diameter" */
inline void cluster_points(const std::vector<Target>& targs, std::vector<std::vector<Target>>& clusters, int bound_pts = 4) {
  clusters.clear();
  std::list<unsigned> pts;
  for (unsigned i = 0; i < targs.size(); ++i) pts.push_back(i);
  std::queue<unsigned> que;
  auto it = pts.begin();
  while (it != pts.end()) {
    clusters.push_back(std::vector<Target>(1, targs[*it]));
    que.push(0);
    pts.erase(it);
    while (!que.empty()) {
      const Target ref = clusters.back()[que.front()];
      que.pop();
      auto it2 = pts.begin();
      while (it2 != pts.end()) {
        const Target& o = targs[*it2];
        // cv::norm(Point2f) accumulates in double; the bound is evaluated in double as well
        const double dx = (double)(ref.u - o.u), dy = (double)(ref.v - o.v);
        if (std::sqrt(dx * dx + dy * dy) <= ((float)bound_pts + 0.5 * (ref.diameter + o.diameter))) {
          que.push((unsigned)clusters.back().size());
          clusters.back().push_back(o);
          it2 = pts.erase(it2);
        } else {
          ++it2;
        }
      }
    }
    it = pts.begin();
  }
}

inline void get_target_boundary(const Target& t, Point2i& t_min, Point2i& t_max) {
  t_min.x = (int)std::floor(t.u - 0.5 * t.diameter);
  t_min.y = (int)std::floor(t.v - 0.5 * t.diameter);
  t_max.x = (int)std::ceil(t.u + 0.5 * t.diameter);
  t_max.y = (int)std::ceil(t.v + 0.5 * t.diameter);
}

inline void get_target_boundary(const Target& t, std::vector<Point2i>& internal, std::vector<Point2i>& bounds,
                                unsigned ubound_pts = 2, unsigned ubuffer = 0) {
  const int bound_pts = (int)ubound_pts, buffer = (int)ubuffer;
  internal.clear();
  bounds.clear();
  Point2i t_min, t_max;
  get_target_boundary(t, t_min, t_max);
  for (int x = t_min.x; x <= t_max.x; ++x)
    for (int y = t_min.y; y <= t_max.y; ++y) internal.push_back({x, y});
  for (int x = t_min.x - bound_pts - buffer; x <= t_max.x + bound_pts + buffer; ++x)
    for (int y = t_min.y - bound_pts - buffer; y <= t_max.y + bound_pts + buffer; ++y)
      if (x < t_min.x - buffer || x > t_max.x + buffer || y < t_min.y - buffer || y > t_max.y + buffer) bounds.push_back({x, y});
}

inline void get_cluster_boundary(const std::vector<Target>& targs, std::vector<Point2i>& internal,
                                 std::vector<Point2i>& bounds, unsigned bound_pts = 2, unsigned buffer = 0) {
  internal.clear();
  bounds.clear();
  std::vector<Point2i> tg_mins(targs.size()), tg_maxs(targs.size());
  Point2i t_max{0, 0}, t_min{std::numeric_limits<int>::max(), std::numeric_limits<int>::max()};
  for (size_t i = 0; i < targs.size(); ++i) {
    get_target_boundary(targs[i], tg_mins[i], tg_maxs[i]);
    t_min.x = std::min(t_min.x, tg_mins[i].x);
    t_min.y = std::min(t_min.y, tg_mins[i].y);
    t_max.x = std::max(t_max.x, tg_maxs[i].x);
    t_max.y = std::max(t_max.y, tg_maxs[i].y);
  }
  const int pad = (int)(bound_pts + buffer);
  t_min.x -= pad; t_min.y -= pad;
  t_max.x += pad; t_max.y += pad;
  const unsigned d_x = (unsigned)(t_max.x - t_min.x + 1), d_y = (unsigned)(t_max.y - t_min.y + 1);
  std::vector<int> grid((size_t)d_x * d_y, 0);
  auto at = [&](unsigned x, unsigned y) -> int& { return grid[(size_t)x * d_y + y]; };
  for (size_t i = 0; i < targs.size(); ++i)
    for (int x = tg_mins[i].x - t_min.x; x <= tg_maxs[i].x - t_min.x; ++x)
      for (int y = tg_mins[i].y - t_min.y; y <= tg_maxs[i].y - t_min.y; ++y) at((unsigned)x, (unsigned)y) = 2;
  // fill between the first and last target pixel of every column, then of every row
  for (unsigned x = 0; x < d_x; ++x) {
    unsigned lo = d_y, hi = d_y;
    for (unsigned y = 0; y < d_y; ++y) if (at(x, y) == 2) { lo = y; break; }
    if (lo == d_y) continue;
    for (unsigned y = d_y - 1;; --y) { if (at(x, y) == 2) { hi = y; break; } if (y == lo) break; }
    for (unsigned y = lo; y <= hi; ++y) at(x, y) = 2;
  }
  for (unsigned y = 0; y < d_y; ++y) {
    unsigned lo = d_x, hi = d_x;
    for (unsigned x = 0; x < d_x; ++x) if (at(x, y) == 2) { lo = x; break; }
    if (lo == d_x) continue;
    for (unsigned x = d_x - 1;; --x) { if (at(x, y) == 2) { hi = x; break; } if (x == lo) break; }
    for (unsigned x = lo; x <= hi; ++x) at(x, y) = 2;
  }
  // maximum of a block [x0, x0+lx) x [y0, y0+ly) (Eigen block().maxCoeff(); the grid changes while the
  // scan runs -- accepted boundary pixels become 1 -- exactly as in the reference)
  auto block_max = [&](unsigned x0, unsigned y0, unsigned lx, unsigned ly) {
    int m = std::numeric_limits<int>::min();
    for (unsigned x = x0; x < x0 + lx; ++x)
      for (unsigned y = y0; y < y0 + ly; ++y) m = std::max(m, at(x, y));
    return m;
  };
  for (unsigned x = 0; x < d_x; ++x) {
    const unsigned min_x = x <= bound_pts + buffer ? 0 : x - bound_pts - buffer;
    const unsigned len_x = std::min(x + bound_pts + buffer, d_x - 1) - min_x + 1;
    const unsigned buf_min_x = x <= buffer ? 0 : x - buffer;
    const unsigned buf_len_x = std::min(x + buffer, d_x - 1) - buf_min_x + 1;
    for (unsigned y = 0; y < d_y; ++y) {
      const unsigned min_y = y <= bound_pts + buffer ? 0 : y - bound_pts - buffer;
      const unsigned len_y = std::min(y + bound_pts + buffer, d_y - 1) - min_y + 1;
      if (at(x, y) == 2) {
        internal.push_back({(int)x + t_min.x, (int)y + t_min.y});
        continue;
      }
      if (bound_pts > 0 && buffer > 0) {
        const unsigned buf_min_y = y <= buffer ? 0 : y - buffer;
        const unsigned buf_len_y = std::min(y + buffer, d_y - 1) - buf_min_y + 1;
        if (block_max(buf_min_x, buf_min_y, buf_len_x, buf_len_y) != 2 && block_max(min_x, min_y, len_x, len_y) == 2) {
          bounds.push_back({(int)x + t_min.x, (int)y + t_min.y});
          at(x, y) = 1;
        }
        continue;
      }
      if (bound_pts > 0 && block_max(min_x, min_y, len_x, len_y) == 2) {
        bounds.push_back({(int)x + t_min.x, (int)y + t_min.y});
        at(x, y) = 1;
      }
    }
  }
}

/* upsp::PatchClusters<float>: the per-cluster pixel lists (the polynomial fit itself runs on the GPU) */
struct PatchClusters {
  int width, height;
  unsigned boundary_thickness, buffer_thickness;
  std::vector<std::vector<unsigned>> bounds_x, bounds_y, internal_x, internal_y;

  PatchClusters(const std::vector<std::vector<Target>>& clusters, int width_in, int height_in,
                unsigned boundary_thickness_in, unsigned buffer_thickness_in)
      : width(width_in), height(height_in), boundary_thickness(boundary_thickness_in), buffer_thickness(buffer_thickness_in),
        bounds_x(clusters.size()), bounds_y(clusters.size()), internal_x(clusters.size()), internal_y(clusters.size()) {
    auto contains = [&](const Point2i& p) { return p.x >= 0 && p.y >= 0 && p.x < width && p.y < height; };   // projection.cpp:10-16
    for (size_t i = 0; i < clusters.size(); ++i) {
      std::vector<Point2i> bounds, internal;
      if (clusters[i].size() > 1) get_cluster_boundary(clusters[i], internal, bounds, boundary_thickness, buffer_thickness);
      else get_target_boundary(clusters[i][0], internal, bounds, boundary_thickness, buffer_thickness);
      for (const auto& p : internal)
        if (contains(p)) { internal_x[i].push_back((unsigned)p.x); internal_y[i].push_back((unsigned)p.y); }
      for (const auto& p : bounds)
        if (contains(p)) { bounds_x[i].push_back((unsigned)p.x); bounds_y[i].push_back((unsigned)p.y); }
    }
  }

  /* ref: [height][width] reference frame (the first frame, psp_process.cpp:2153-2163: thresh from its histogram, offset 2) */
  template <typename T>
  void threshold_bounds(const T* ref, unsigned thresh, unsigned offset) {
    const int off = (int)offset, col_idxs = width - 1, row_idxs = height - 1;
    for (size_t i = 0; i < bounds_x.size(); ++i) {
      std::vector<unsigned> kx, ky;
      for (size_t j = 0; j < bounds_x[i].size(); ++j) {
        const int x = (int)bounds_x[i][j], y = (int)bounds_y[i][j];
        const int y_min = std::max(0, y - off), x_min = std::max(0, x - off);
        const int w = std::min(col_idxs, x + off) - x_min + 1, h = std::min(row_idxs, y + off) - y_min + 1;
        double min_val = std::numeric_limits<double>::max();
        for (int yy = y_min; yy < y_min + h; ++yy)
          for (int xx = x_min; xx < x_min + w; ++xx) min_val = std::min(min_val, (double)ref[(size_t)yy * width + xx]);
        if (!(min_val < thresh)) { kx.push_back(bounds_x[i][j]); ky.push_back(bounds_y[i][j]); }
      }
      bounds_x[i].swap(kx);
      bounds_y[i].swap(ky);
    }
  }

  /* the flattened form of upsp_gpu_set_patches: offsets [n_clusters+1] and concatenated coordinates */
  void flatten(std::vector<int32_t>& boff, std::vector<uint32_t>& bx, std::vector<uint32_t>& by,
               std::vector<int32_t>& ioff, std::vector<uint32_t>& ix, std::vector<uint32_t>& iy) const {
    boff.assign(1, 0); ioff.assign(1, 0);
    bx.clear(); by.clear(); ix.clear(); iy.clear();
    for (size_t i = 0; i < bounds_x.size(); ++i) {
      bx.insert(bx.end(), bounds_x[i].begin(), bounds_x[i].end());
      by.insert(by.end(), bounds_y[i].begin(), bounds_y[i].end());
      ix.insert(ix.end(), internal_x[i].begin(), internal_x[i].end());
      iy.insert(iy.end(), internal_y[i].begin(), internal_y[i].end());
      boff.push_back((int32_t)bx.size());
      ioff.push_back((int32_t)ix.size());
    }
  }
};

}  // namespace upsp_b200
