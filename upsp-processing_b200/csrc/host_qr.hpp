// host_qr.hpp -- host-side, frame-invariant factorisation used by the fiducial patcher.
//
// polyfit2D (cpp/lib/patches.ipp:172-204) solves A p = z with
// A.colPivHouseholderQr().solve(z) in float, A being the cubic 2-D Vandermonde of the
// cluster's boundary pixels.  A depends on geometry only, so the factorisation is done once
// per cluster here and the per-frame kernel (k_patch) only applies it.  The sequence below
// is Eigen 3.4's ColPivHouseholderQR::computeInPlace / makeHouseholder /
// applyHouseholderOnTheLeft with scalar left-to-right reductions, in float, so that the
// stored reflectors are the ones the reference would apply.  Compile with
// -ffp-contract=off (no FMA): the system is numerically rank deficient in float and the
// result is only reproducible operation for operation.
#pragma once
#include <cfloat>
#include <cmath>
#include <utility>
#include <vector>

namespace upsp {

struct ColPivQR {
  int rows = 0, cols = 0, nonzero_pivots = 0;
  std::vector<float> qr;     // col-major rows x cols: R on/above the diagonal, essential parts below
  std::vector<float> hcoef;  // tau_k
  std::vector<int> perm;     // column permutation indices
};

inline float sq_norm(const float* v, int n) {
  float s = 0.0f;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return s;
}

inline ColPivQR colpiv_householder_qr(std::vector<float> a, int rows, int cols) {
  ColPivQR out;
  out.rows = rows;
  out.cols = cols;
  const int size = rows < cols ? rows : cols;
  std::vector<float> upd(cols), dir(cols), tmp(cols), hc(size, 0.0f);
  std::vector<int> transp(size);
  auto colp = [&](int j) { return a.data() + (size_t)j * rows; };
  float maxnorm = 0.0f;
  for (int k = 0; k < cols; ++k) {
    dir[k] = std::sqrt(sq_norm(colp(k), rows));
    upd[k] = dir[k];
    if (k == 0 || upd[k] > maxnorm) maxnorm = upd[k];
  }
  const float me = maxnorm * FLT_EPSILON;
  const float threshold_helper = (me * me) / (float)rows;
  const float downdate_threshold = std::sqrt(FLT_EPSILON);
  int nzp = size;
  for (int k = 0; k < size; ++k) {
    int big = k;
    for (int j = k + 1; j < cols; ++j)
      if (upd[j] > upd[big]) big = j;
    const float big_sq = upd[big] * upd[big];
    if (nzp == size && big_sq < threshold_helper * (float)(rows - k)) nzp = k;
    transp[k] = big;
    if (big != k) {
      float *p = colp(k), *q = colp(big);
      for (int i = 0; i < rows; ++i) std::swap(p[i], q[i]);
      std::swap(upd[k], upd[big]);
      std::swap(dir[k], dir[big]);
    }
    float* v = colp(k) + k;
    const int n = rows - k;
    const float tail = n == 1 ? 0.0f : sq_norm(v + 1, n - 1);
    const float c0 = v[0];
    float beta, tau;
    if (tail <= FLT_MIN) {
      tau = 0.0f;
      beta = c0;
      for (int i = 1; i < n; ++i) v[i] = 0.0f;
    } else {
      beta = std::sqrt(c0 * c0 + tail);
      if (c0 >= 0.0f) beta = -beta;
      const float den = c0 - beta;
      for (int i = 1; i < n; ++i) v[i] = v[i] / den;
      tau = (beta - c0) / beta;
    }
    hc[k] = tau;
    v[0] = beta;
    const int nc = cols - k - 1;
    if (nc > 0) {
      if (n == 1) {
        for (int j = 0; j < nc; ++j) colp(k + 1 + j)[k] *= (1.0f - tau);
      } else if (tau != 0.0f) {
        const float* e = v + 1;
        for (int j = 0; j < nc; ++j) {
          const float* b = colp(k + 1 + j) + k;
          float s = 0.0f;
          for (int i = 0; i < n - 1; ++i) s += e[i] * b[1 + i];
          tmp[j] = s + b[0];
        }
        for (int j = 0; j < nc; ++j) {
          float* b = colp(k + 1 + j) + k;
          b[0] -= tau * tmp[j];
          for (int i = 0; i < n - 1; ++i) b[1 + i] -= tmp[j] * (tau * e[i]);
        }
      }
    }
    for (int j = k + 1; j < cols; ++j) {
      if (upd[j] != 0.0f) {
        float t = std::fabs(colp(j)[k]) / upd[j];
        t = (1.0f + t) * (1.0f - t);
        t = t < 0.0f ? 0.0f : t;
        const float r = upd[j] / dir[j];
        const float t2 = t * (r * r);
        if (t2 <= downdate_threshold) {
          dir[j] = std::sqrt(sq_norm(colp(j) + k + 1, rows - k - 1));
          upd[j] = dir[j];
        } else {
          upd[j] *= std::sqrt(t);
        }
      }
    }
  }
  out.perm.resize(cols);
  for (int i = 0; i < cols; ++i) out.perm[i] = i;
  for (int k = 0; k < size; ++k) std::swap(out.perm[k], out.perm[transp[k]]);
  out.nonzero_pivots = nzp;
  out.hcoef = std::move(hc);
  out.qr = std::move(a);
  return out;
}

}  // namespace upsp
