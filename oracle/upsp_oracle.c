/*
 * upsp_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's (nasa/upsp-processing) per-frame
 * psp_process chain, used ONLY as the checker in tests/, in
 * __graft_entry__.smoke() and as bench.py's cpu_baseline / --impl reference
 * arm.  Nothing under upsp-processing_b200/ may include, link or call it; the
 * product path fails loudly if its CUDA library is missing.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).  Arithmetic that the reference delegates to un-vendored
 * third parties is restated from their published algorithms:
 *   - OpenCV 4.x cv::warpAffine (fixed-point 1/32-px bilinear; imgwarp.cpp) --
 *     PINNED against cv2.warpAffine golden vectors (tests/golden/warp_*.npz).
 *   - Eigen 3.4 ColPivHouseholderQR (compute + solve) in float, scalar
 *     left-to-right reduction order.  PARITY UNPINNED: Eigen is not in this
 *     container and its packet reductions depend on build flags; see DESIGN.md.
 *   - Eigen row-major SparseMatrix * dense vector (single accumulator).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off because the reference's default x86-64 build has no FMA.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* a1: video decode.  cpp/lib/PSPVideo.cpp:134-150 (unpack_12bit),          */
/*     :111-132 (unpack_10bit), 10->12 bit LUT cpp/lib/CineReader.cpp:409-425 */
/* ------------------------------------------------------------------------ */
ORC_API void orc_unpack_12bit(const uint8_t* packed, size_t nbytes, uint16_t* dst) {
  for (size_t i = 0; i + 3 <= nbytes; i += 3, dst += 2) {
    uint16_t p = packed[i], q = packed[i + 1], r = packed[i + 2];
    dst[0] = (uint16_t)((p << 4) | (q >> 4));
    dst[1] = (uint16_t)(((q & 0xF) << 8) | r);
  }
}

/* a batch of frames, one frame per loop iteration (the reference's ranks each decode their own frames,
 * cpp/exec/psp_process.cpp:1766-1773): used by bench.py's CPU arm so that it starts from packed bytes */
ORC_API void orc_unpack_12bit_frames(const uint8_t* packed, size_t frame_bytes, int n_frames, uint16_t* dst) {
  const size_t npix = frame_bytes / 3 * 2;
#pragma omp parallel for schedule(static)
  for (int f = 0; f < n_frames; ++f) orc_unpack_12bit(packed + (size_t)f * frame_bytes, frame_bytes, dst + (size_t)f * npix);
}

ORC_API void orc_unpack_10bit(const uint8_t* packed, size_t nbytes, uint16_t* dst,
                              const uint16_t* lut /* 1024 entries or NULL */) {
  for (size_t i = 0; i + 5 <= nbytes; i += 5, dst += 4) {
    uint16_t p = packed[i], q = packed[i + 1], r = packed[i + 2], s = packed[i + 3],
             t = packed[i + 4];
    dst[0] = (uint16_t)((p << 2) | (q >> 6));
    dst[1] = (uint16_t)(((q & 0x3F) << 4) | (r >> 4));
    dst[2] = (uint16_t)(((r & 0x0F) << 6) | (s >> 2));
    dst[3] = (uint16_t)(((s & 0x03) << 8) | t);
    if (lut)
      for (int k = 0; k < 4; ++k) dst[k] = lut[dst[k]];
  }
}

/* ------------------------------------------------------------------------ */
/* a2: hot-pixel fix.  cpp/utils/cv_extras.cpp:230-272; defaults            */
/*     cpp/include/utils/cv_extras.h:154-155 (4064, 512, 5).                */
/* returns number of hot pixels found, or -1 if more than max_hot (no-op).  */
/* ------------------------------------------------------------------------ */
static int cmp_u16(const void* a, const void* b) {
  return (int)*(const uint16_t*)a - (int)*(const uint16_t*)b;
}

ORC_API int orc_fix_hot_pixels(uint16_t* img, int rows, int cols, int thresh,
                               int min_change, int max_hot) {
  int n_pix = rows * cols;
  int* locs = (int*)malloc(sizeof(int) * (size_t)(max_hot > 0 ? max_hot : 1));
  int n_hot = 0;
  for (int pix = 0; pix < n_pix; ++pix) {
    if (img[pix] >= thresh) {
      if (n_hot >= max_hot) {
        free(locs);
        return -1;
      }
      locs[n_hot++] = pix;
    }
  }
  for (int h = 0; h < n_hot; ++h) {
    uint16_t vals[4];
    int n_vals = 0;
    int row = locs[h] / cols, col = locs[h] % cols;
    if (row > 0) vals[n_vals++] = img[(row - 1) * cols + col];
    if (col > 0) vals[n_vals++] = img[row * cols + col - 1];
    if (row < rows - 1) vals[n_vals++] = img[(row + 1) * cols + col];
    if (col < cols - 1) vals[n_vals++] = img[row * cols + col + 1];
    qsort(vals, (size_t)n_vals, sizeof(uint16_t), cmp_u16);
    uint16_t old_val = img[row * cols + col];
    uint16_t new_val = vals[n_vals / 2];
    if ((int)old_val - (int)new_val > min_change) img[row * cols + col] = new_val;
  }
  free(locs);
  return n_hot;
}

/* ------------------------------------------------------------------------ */
/* a3: cv::warpAffine(..., flags | WARP_INVERSE_MAP), BORDER_CONSTANT 0,    */
/*     as called at cpp/lib/registration.cpp:69-72.                          */
/* OpenCV model (imgwarp.cpp WarpAffineInvoker + remapBilinear/remapNearest):*/
/*   AB_BITS=10, INTER_BITS=5; coordinates in int32 fixed point, weights    */
/*   (1-fy)(1-fx).. as float, sum left-to-right in float, cvRound+saturate.  */
/* interp: 0 = INTER_NEAREST, 1 = INTER_LINEAR                               */
/* ------------------------------------------------------------------------ */
static inline int orc_cvround(double v) { return (int)lrint(v); }

static inline void warp_coords(const double M[6], int x, int y, int round_delta,
                               int* X, int* Y) {
  int adelta = orc_cvround(M[0] * x * 1024.0);
  int bdelta = orc_cvround(M[3] * x * 1024.0);
  int X0 = orc_cvround((M[1] * y + M[2]) * 1024.0) + round_delta;
  int Y0 = orc_cvround((M[4] * y + M[5]) * 1024.0) + round_delta;
  *X = X0 + adelta;
  *Y = Y0 + bdelta;
}

static inline uint16_t sat_u16(float v) {
  int iv = (int)lrintf(v);
  return (uint16_t)(iv < 0 ? 0 : (iv > 65535 ? 65535 : iv));
}

ORC_API void orc_warp_affine_u16(const uint16_t* src, int sw, int sh, const float Mf[6],
                                 int interp, uint16_t* dst, int dw, int dh) {
  double M[6];
  for (int i = 0; i < 6; ++i) M[i] = (double)Mf[i];
  /* per-column tables, exactly WarpAffineInvoker's adelta / bdelta */
  int* adelta = (int*)malloc(sizeof(int) * 2 * (size_t)dw);
  int* bdelta = adelta + dw;
  for (int x = 0; x < dw; ++x) {
    adelta[x] = orc_cvround(M[0] * x * 1024.0);
    bdelta[x] = orc_cvround(M[3] * x * 1024.0);
  }
  const int round_delta = interp == 0 ? 512 : 16;
  for (int y = 0; y < dh; ++y) {
    const int X0 = orc_cvround((M[1] * y + M[2]) * 1024.0) + round_delta;
    const int Y0 = orc_cvround((M[4] * y + M[5]) * 1024.0) + round_delta;
    uint16_t* drow = dst + (size_t)y * dw;
    for (int x = 0; x < dw; ++x) {
      int X = X0 + adelta[x], Y = Y0 + bdelta[x];
      if (interp == 0) {
        int sx = X >> 10, sy = Y >> 10;
        drow[x] = (sx >= 0 && sx < sw && sy >= 0 && sy < sh) ? src[(size_t)sy * sw + sx] : 0;
      } else {
        X >>= 5;
        Y >>= 5;
        int sx = X >> 5, sy = Y >> 5;
        float fx = (float)(X & 31) / 32.0f, fy = (float)(Y & 31) / 32.0f;
        float w0 = (1.0f - fy) * (1.0f - fx), w1 = (1.0f - fy) * fx;
        float w2 = fy * (1.0f - fx), w3 = fy * fx;
        if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
          const uint16_t* s0 = src + (size_t)sy * sw + sx;
          float sum = (float)s0[0] * w0 + (float)s0[1] * w1 + (float)s0[sw] * w2 + (float)s0[sw + 1] * w3;
          drow[x] = sat_u16(sum);
          continue;
        }
        if (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0) {
          drow[x] = 0;
          continue;
        }
        int x0 = sx >= 0 && sx < sw, x1 = sx + 1 >= 0 && sx + 1 < sw;
        int y0 = sy >= 0 && sy < sh, y1 = sy + 1 >= 0 && sy + 1 < sh;
        float v0 = (x0 && y0) ? (float)src[(size_t)sy * sw + sx] : 0.0f;
        float v1 = (x1 && y0) ? (float)src[(size_t)sy * sw + sx + 1] : 0.0f;
        float v2 = (x0 && y1) ? (float)src[(size_t)(sy + 1) * sw + sx] : 0.0f;
        float v3 = (x1 && y1) ? (float)src[(size_t)(sy + 1) * sw + sx + 1] : 0.0f;
        float sum = v0 * w0 + v1 * w1 + v2 * w2 + v3 * w3;
        drow[x] = sat_u16(sum);
      }
    }
  }
  free(adelta);
}

/* f32 flavour (used by the ECC restatement: cv::findTransformECC warps the   */
/* f32 image and its gradients with the same fixed-point coordinates).        */
ORC_API void orc_warp_affine_f32(const float* src, int sw, int sh, const float Mf[6],
                                 int interp, float* dst, int dw, int dh) {
  double M[6];
  for (int i = 0; i < 6; ++i) M[i] = (double)Mf[i];
  for (int y = 0; y < dh; ++y) {
    for (int x = 0; x < dw; ++x) {
      int X, Y;
      if (interp == 0) {
        warp_coords(M, x, y, 512, &X, &Y);
        int sx = X >> 10, sy = Y >> 10;
        dst[(size_t)y * dw + x] =
            (sx >= 0 && sx < sw && sy >= 0 && sy < sh) ? src[(size_t)sy * sw + sx] : 0.0f;
      } else {
        warp_coords(M, x, y, 16, &X, &Y);
        X >>= 5;
        Y >>= 5;
        int sx = X >> 5, sy = Y >> 5;
        float fx = (float)(X & 31) / 32.0f, fy = (float)(Y & 31) / 32.0f;
        float w0 = (1.0f - fy) * (1.0f - fx), w1 = (1.0f - fy) * fx;
        float w2 = fy * (1.0f - fx), w3 = fy * fx;
        if (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0) {
          dst[(size_t)y * dw + x] = 0.0f;
          continue;
        }
        int x0 = sx >= 0 && sx < sw, x1 = sx + 1 >= 0 && sx + 1 < sw;
        int y0 = sy >= 0 && sy < sh, y1 = sy + 1 >= 0 && sy + 1 < sh;
        float v0 = (x0 && y0) ? src[(size_t)sy * sw + sx] : 0.0f;
        float v1 = (x1 && y0) ? src[(size_t)sy * sw + sx + 1] : 0.0f;
        float v2 = (x0 && y1) ? src[(size_t)(sy + 1) * sw + sx] : 0.0f;
        float v3 = (x1 && y1) ? src[(size_t)(sy + 1) * sw + sx + 1] : 0.0f;
        dst[(size_t)y * dw + x] = v0 * w0 + v1 * w1 + v2 * w2 + v3 * w3;
      }
    }
  }
}

/* ------------------------------------------------------------------------ */
/* Eigen::ColPivHouseholderQR<float> restated (Eigen 3.4                      */
/* ColPivHouseholderQR.h computeInPlace / _solve_impl, Householder.h          */
/* makeHouseholder / applyHouseholderOnTheLeft).  Used by                     */
/* cpp/lib/patches.ipp:203 and cpp/lib/filtering.ipp:65.  Column-major.      */
/* Reductions are scalar left-to-right (Eigen's packet order is build-flag    */
/* dependent: PARITY UNPINNED for the last bits).                             */
/* ------------------------------------------------------------------------ */
static float vec_sqnorm(const float* a, int n) {
  float s = 0.0f;
  for (int i = 0; i < n; ++i) s += a[i] * a[i];
  return s;
}

/* A: rows x cols col-major, overwritten by the factorisation.
 * hcoef[cols], transp[cols] (column transpositions), returns nonzero_pivots. */
ORC_API int orc_colpiv_qr_f32(float* A, int rows, int cols, float* hcoef, int* transp) {
  int size = rows < cols ? rows : cols;
  float* norms_upd = (float*)malloc(sizeof(float) * (size_t)cols);
  float* norms_dir = (float*)malloc(sizeof(float) * (size_t)cols);
  float* tmp = (float*)malloc(sizeof(float) * (size_t)cols);
  float maxnorm = 0.0f;
  for (int k = 0; k < cols; ++k) {
    norms_dir[k] = sqrtf(vec_sqnorm(A + (size_t)k * rows, rows));
    norms_upd[k] = norms_dir[k];
    if (k == 0 || norms_upd[k] > maxnorm) maxnorm = norms_upd[k];
  }
  float th = maxnorm * FLT_EPSILON;
  float threshold_helper = (th * th) / (float)rows;
  float norm_downdate_threshold = sqrtf(FLT_EPSILON);
  int nonzero_pivots = size;

  for (int k = 0; k < size; ++k) {
    int big = k;
    float bigv = norms_upd[k];
    for (int j = k + 1; j < cols; ++j)
      if (norms_upd[j] > bigv) {
        bigv = norms_upd[j];
        big = j;
      }
    float big_sq = bigv * bigv;
    if (nonzero_pivots == size && big_sq < threshold_helper * (float)(rows - k))
      nonzero_pivots = k;
    transp[k] = big;
    if (k != big) {
      float* ck = A + (size_t)k * rows;
      float* cb = A + (size_t)big * rows;
      for (int i = 0; i < rows; ++i) {
        float t = ck[i];
        ck[i] = cb[i];
        cb[i] = t;
      }
      float t = norms_upd[k];
      norms_upd[k] = norms_upd[big];
      norms_upd[big] = t;
      t = norms_dir[k];
      norms_dir[k] = norms_dir[big];
      norms_dir[big] = t;
    }
    /* makeHouseholderInPlace on col(k).tail(rows-k) */
    float* v = A + (size_t)k * rows + k;
    int n = rows - k;
    float tail_sq = (n == 1) ? 0.0f : vec_sqnorm(v + 1, n - 1);
    float c0 = v[0], beta, tau;
    if (tail_sq <= FLT_MIN) {
      tau = 0.0f;
      beta = c0;
      for (int i = 1; i < n; ++i) v[i] = 0.0f;
    } else {
      beta = sqrtf(c0 * c0 + tail_sq);
      if (c0 >= 0.0f) beta = -beta;
      float den = c0 - beta;
      for (int i = 1; i < n; ++i) v[i] = v[i] / den;
      tau = (beta - c0) / beta;
    }
    hcoef[k] = tau;
    v[0] = beta;
    /* apply H_k to bottomRightCorner(rows-k, cols-k-1) */
    int nc = cols - k - 1;
    if (nc > 0) {
      if (n == 1) {
        for (int j = 0; j < nc; ++j) A[(size_t)(k + 1 + j) * rows + k] *= (1.0f - tau);
      } else if (tau != 0.0f) {
        const float* e = v + 1;
        for (int j = 0; j < nc; ++j) {
          float* b = A + (size_t)(k + 1 + j) * rows + k;
          float s = 0.0f;
          for (int i = 0; i < n - 1; ++i) s += e[i] * b[1 + i];
          tmp[j] = s + b[0];
        }
        for (int j = 0; j < nc; ++j) {
          float* b = A + (size_t)(k + 1 + j) * rows + k;
          b[0] -= tau * tmp[j];
          for (int i = 0; i < n - 1; ++i) b[1 + i] -= tmp[j] * (tau * e[i]);
        }
      }
    }
    /* norm downdate (LAPACK lawn176 as in Eigen) */
    for (int j = k + 1; j < cols; ++j) {
      if (norms_upd[j] != 0.0f) {
        float t = fabsf(A[(size_t)j * rows + k]) / norms_upd[j];
        t = (1.0f + t) * (1.0f - t);
        t = t < 0.0f ? 0.0f : t;
        float r = norms_upd[j] / norms_dir[j];
        float t2 = t * (r * r);
        if (t2 <= norm_downdate_threshold) {
          norms_dir[j] = sqrtf(vec_sqnorm(A + (size_t)j * rows + k + 1, rows - k - 1));
          norms_upd[j] = norms_dir[j];
        } else {
          norms_upd[j] *= sqrtf(t);
        }
      }
    }
  }
  free(norms_upd);
  free(norms_dir);
  free(tmp);
  return nonzero_pivots;
}

/* x[cols] = solve(QR, b[rows]); c is scratch of length rows. */
ORC_API void orc_colpiv_qr_solve_f32(const float* QR, int rows, int cols, const float* hcoef,
                                     const int* transp, int nonzero_pivots, const float* b,
                                     float* x, float* c) {
  int size = rows < cols ? rows : cols;
  if (nonzero_pivots == 0) {
    for (int i = 0; i < cols; ++i) x[i] = 0.0f;
    return;
  }
  memcpy(c, b, sizeof(float) * (size_t)rows);
  /* c = Q^T c : apply H_0, H_1, ... H_{nz-1} */
  for (int k = 0; k < nonzero_pivots; ++k) {
    int n = rows - k;
    float tau = hcoef[k];
    if (n == 1) {
      c[k] *= (1.0f - tau);
    } else if (tau != 0.0f) {
      const float* e = QR + (size_t)k * rows + k + 1;
      float s = 0.0f;
      for (int i = 0; i < n - 1; ++i) s += e[i] * c[k + 1 + i];
      float t = s + c[k];
      c[k] -= tau * t;
      for (int i = 0; i < n - 1; ++i) c[k + 1 + i] -= t * (tau * e[i]);
    }
  }
  /* R[0:nz,0:nz] upper-triangular back substitution, column oriented */
  for (int i = nonzero_pivots - 1; i >= 0; --i) {
    c[i] = c[i] / QR[(size_t)i * rows + i];
    for (int j = 0; j < i; ++j) c[j] -= c[i] * QR[(size_t)i * rows + j];
  }
  /* column permutation from the transpositions */
  int* perm = (int*)malloc(sizeof(int) * (size_t)cols);
  for (int i = 0; i < cols; ++i) perm[i] = i;
  for (int k = 0; k < size; ++k) {
    int t = perm[k];
    perm[k] = perm[transp[k]];
    perm[transp[k]] = t;
  }
  for (int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = c[i];
  for (int i = nonzero_pivots; i < cols; ++i) x[perm[i]] = 0.0f;
  free(perm);
}

/* ------------------------------------------------------------------------ */
/* a4: PatchClusters<float>::operator() cpp/lib/patches.ipp:98-164,          */
/*     polyfit2D :172-204, polyval2D :208-236.  img is f32 [H*W], in place.  */
/* Clusters are applied in order (later clusters see earlier patches).       */
/* ------------------------------------------------------------------------ */
ORC_API void orc_patch_apply(float* img, int width, int n_clusters, const int* bounds_off,
                             const uint32_t* bx, const uint32_t* by, const int* internal_off,
                             const uint32_t* ix, const uint32_t* iy) {
  const int degree = 3, coeffs = 10;
  for (int cl = 0; cl < n_clusters; ++cl) {
    int nb = bounds_off[cl + 1] - bounds_off[cl];
    if (nb < coeffs) continue;
    const uint32_t* x = bx + bounds_off[cl];
    const uint32_t* y = by + bounds_off[cl];
    float* A = (float*)malloc(sizeof(float) * (size_t)nb * coeffs);
    float* z = (float*)malloc(sizeof(float) * (size_t)nb);
    float* c = (float*)malloc(sizeof(float) * (size_t)nb);
    for (int ind = 0; ind < nb; ++ind) {
      z[ind] = img[(size_t)y[ind] * width + x[ind]];
      int count = 0;
      for (int i = 0; i <= degree; ++i)
        for (int j = 0; j <= degree; ++j)
          if (i + j <= degree) {
            A[(size_t)count * nb + ind] =
                (float)pow((double)y[ind], i) * (float)pow((double)x[ind], j);
            ++count;
          }
    }
    float hcoef[10], poly[10];
    int transp[10];
    int nz = orc_colpiv_qr_f32(A, nb, coeffs, hcoef, transp);
    orc_colpiv_qr_solve_f32(A, nb, coeffs, hcoef, transp, nz, z, poly, c);
    int ni = internal_off[cl + 1] - internal_off[cl];
    const uint32_t* px = ix + internal_off[cl];
    const uint32_t* py = iy + internal_off[cl];
    for (int ind = 0; ind < ni; ++ind) {
      float zz = 0.0f;
      int count = 0;
      for (int i = 0; i <= degree; ++i)
        for (int j = 0; j <= degree; ++j)
          if (i + j <= degree) {
            zz += poly[count] * (float)pow((double)py[ind], i) * (float)pow((double)px[ind], j);
            ++count;
          }
      img[(size_t)py[ind] * width + px[ind]] = zz;
    }
    free(A);
    free(z);
    free(c);
  }
}

/* ------------------------------------------------------------------------ */
/* a5: optional spatial filter cpp/exec/psp_process.cpp:1802-1807:            */
/*   cv::GaussianBlur(img, img, Size(k,k), 0) | cv::blur(img, img, Size(k,k)) */
/* default border BORDER_REFLECT_101.  kind: 1 gaussian, 2 box.               */
/* CV_16U gaussian = OpenCV's fixed-point path (16 fractional bits per pass,  */
/* exact accumulation, one rounding) -- PINNED against cv2 goldens for        */
/* k = 3,5,7 (sigma = 0 -> getGaussianKernel's fixed small kernels) and, with */
/* taps from orc_set_gauss_taps, for every odd size up to 31;                 */
/* CV_16U box = integer sum * (1/area) in double, cvRound; CV_32F gaussian =  */
/* separable float filter in OpenCV's symmetric order (exact on 12-bit data,  */
/* last-bit differences possible next to patched pixels: cv2 may use FMA);    */
/* CV_32F box = double window sum * (1/area).                                 */
/* ------------------------------------------------------------------------ */
static inline int orc_reflect101(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}
static const double* orc_small_gauss(int k) {
  static const double k3[] = {0.25, 0.5, 0.25};
  static const double k5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625};
  static const double k7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
  return k == 3 ? k3 : (k == 5 ? k5 : (k == 7 ? k7 : NULL));
}

/* 16.16 fixed-point taps of sizes beyond 7: supplied by the caller (oracle.py derives them from cv2's bit-exact
 * sigma = 0 kernel with OpenCV's error-diffusion rounding, getGaussianKernelFixedPoint_ED) */
static long long orc_gauss_q[32][31];
static int orc_gauss_q_set[32];
ORC_API void orc_set_gauss_taps(int ksize, const int* q) {
  if (ksize < 1 || ksize > 31) return;
  for (int i = 0; i < ksize; ++i) orc_gauss_q[ksize][i] = q[i];
  orc_gauss_q_set[ksize] = 1;
}

ORC_API int orc_filter_u16(const uint16_t* src, uint16_t* dst, int W, int H, int kind, int ksize) {
  const int r = ksize / 2;
  const double* kd = orc_small_gauss(ksize);
  long long q[31];
  if (kind == 1) {
    if (ksize == 1) q[0] = 65536;
    else if (kd) for (int i = 0; i < ksize; ++i) q[i] = (long long)llrint(kd[i] * 65536.0);
    else if (ksize >= 1 && ksize <= 31 && orc_gauss_q_set[ksize]) for (int i = 0; i < ksize; ++i) q[i] = orc_gauss_q[ksize][i];
    else return 1;
  }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      if (kind == 1) {
        long long S = 0;
        for (int j = -r; j <= r; ++j) {
          long long t = 0;
          for (int i = -r; i <= r; ++i)
            t += q[i + r] * src[(size_t)orc_reflect101(y + j, H) * W + orc_reflect101(x + i, W)];
          S += q[j + r] * t;
        }
        long long v = (S + (1LL << 31)) >> 32;
        dst[(size_t)y * W + x] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
      } else {
        int S = 0;
        for (int j = -r; j <= r; ++j)
          for (int i = -r; i <= r; ++i) S += src[(size_t)orc_reflect101(y + j, H) * W + orc_reflect101(x + i, W)];
        long v = lrint((double)S * (1.0 / (double)(ksize * ksize)));
        dst[(size_t)y * W + x] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
      }
    }
  return 0;
}

ORC_API int orc_filter_f32(const float* src, float* dst, int W, int H, int kind, int ksize) {
  const int r = ksize / 2;
  const double* kd = orc_small_gauss(ksize);
  if (ksize == 1) {   /* a 1x1 window is the identity for both kinds */
    memcpy(dst, src, sizeof(float) * (size_t)W * H);
    return 0;
  }
  if (kind == 1 && !kd) return 1;
  if (kind == 1) {
    float* tmp = (float*)malloc(sizeof(float) * (size_t)W * H);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float* row = src + (size_t)y * W;
        float s = row[x] * (float)kd[r];
        for (int i = 1; i <= r; ++i) s = s + (row[orc_reflect101(x - i, W)] + row[orc_reflect101(x + i, W)]) * (float)kd[r + i];
        tmp[(size_t)y * W + x] = s;
      }
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        float s = tmp[(size_t)y * W + x] * (float)kd[r];
        for (int i = 1; i <= r; ++i)
          s = s + (tmp[(size_t)orc_reflect101(y - i, H) * W + x] + tmp[(size_t)orc_reflect101(y + i, H) * W + x]) * (float)kd[r + i];
        dst[(size_t)y * W + x] = s;
      }
    free(tmp);
  } else {
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        double S = 0.0;
        for (int j = -r; j <= r; ++j) {
          double t = 0.0;
          for (int i = -r; i <= r; ++i) t += (double)src[(size_t)orc_reflect101(y + j, H) * W + orc_reflect101(x + i, W)];
          S += t;
        }
        dst[(size_t)y * W + x] = (float)(S * (1.0 / (double)(ksize * ksize)));
      }
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* a6: upsp::project_frame cpp/lib/projection.ipp:884-908 -- Eigen row-major */
/*     CSR * dense vector, out zero-initialised, one accumulator per row.    */
/* ------------------------------------------------------------------------ */
ORC_API void orc_project_frame(const int* rowptr, const int* col, const float* val, int n_rows,
                               const float* frame, float* out) {
  for (int r = 0; r < n_rows; ++r) {
    float t = 0.0f;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) t += val[k] * frame[col[k]];
    out[r] = 0.0f + 1.0f * t;
  }
}

/* ------------------------------------------------------------------------ */
/* apportion cpp/exec/psp_process.cpp:611-624                                */
/* ------------------------------------------------------------------------ */
ORC_API void orc_apportion(int value, int n_bins, int* start, int* extent) {
  unsigned long block = (unsigned long)(value / n_bins);
  unsigned long rem = (unsigned long)value - block * (unsigned long)n_bins;
  unsigned long next = 0;
  for (unsigned long b = 0; b < (unsigned long)n_bins; ++b) {
    start[b] = (int)next;
    extent[b] = (int)(block + (b < rem));
    next += (unsigned long)extent[b];
  }
}

/* ------------------------------------------------------------------------ */
/* a3-a9: the phase-1 frame loop cpp/exec/psp_process.cpp:1743-1851.         */
/*  frames[c]   : u16 [n_frames][h[c]*w[c]] (this rank's slice)              */
/*  warp[c]     : float [n_frames][6] or NULL (registration result supplied  */
/*                by the caller; the ECC solve lives in oracle.py via cv2)   */
/*  patch arrays: per camera, or n_clusters[c]==0                             */
/*  skipped     : node list (identify_skipped_nodes projection.ipp:858-880)  */
/*  remap       : src_index[N] static form of P3DModel adjust_solution       */
/*                (cpp/lib/P3DModel.ipp:144-157) or NULL                      */
/*  outputs     : intensity [n_frames][N] f32; sum/sumsq [N] f64 (+=)        */
/* ------------------------------------------------------------------------ */
typedef struct {
  int n_cams;
  int n_nodes;
  int n_frames;      /* local */
  int first_frame;   /* global index of local frame 0 (frame 0 is never registered, :1777) */
  const int* width;  /* [n_cams] */
  const int* height;
  const uint16_t* const* frames;
  const float* const* warp;
  int interp;        /* 0 nearest, 1 linear */
  int hot_pixel_fix; /* reference: always 1 */
  const int* const* rowptr;
  const int* const* col;
  const float* const* val;
  const int* n_clusters;
  const int* const* bounds_off;
  const uint32_t* const* bx;
  const uint32_t* const* by;
  const int* const* internal_off;
  const uint32_t* const* ix;
  const uint32_t* const* iy;
  int n_skipped;
  const int* skipped;
  const int* remap;
  int filter_kind;   /* 0 none, 1 gaussian, 2 box (psp_process.cpp:1802-1807) */
  int filter_size;
} orc_phase1_args;

ORC_API void orc_phase1(const orc_phase1_args* a, float* intensity, double* sum, double* sumsq) {
  const int N = a->n_nodes;
#pragma omp parallel
  {
    double* lsum = (double*)calloc((size_t)N, sizeof(double));
    double* lsq = (double*)calloc((size_t)N, sizeof(double));
    float* sol = (float*)malloc(sizeof(float) * (size_t)N);
    float* csol = (float*)malloc(sizeof(float) * (size_t)N);
    float* tmpsol = a->remap ? (float*)malloc(sizeof(float) * (size_t)N) : NULL;
    size_t maxpix = 0;
    for (int c = 0; c < a->n_cams; ++c) {
      size_t p = (size_t)a->width[c] * a->height[c];
      if (p > maxpix) maxpix = p;
    }
    uint16_t* img16 = (uint16_t*)malloc(sizeof(uint16_t) * maxpix);
    uint16_t* warp16 = (uint16_t*)malloc(sizeof(uint16_t) * maxpix);
    float* img32 = (float*)malloc(sizeof(float) * maxpix);
#pragma omp for schedule(dynamic, 1) nowait
    for (int off = 0; off < a->n_frames; ++off) {
      int f = a->first_frame + off;
      for (int c = 0; c < a->n_cams; ++c) {
        int w = a->width[c], h = a->height[c];
        size_t P = (size_t)w * h;
        memcpy(img16, a->frames[c] + (size_t)off * P, P * sizeof(uint16_t));
        if (a->hot_pixel_fix) orc_fix_hot_pixels(img16, h, w, 4064, 512, 5);
        const uint16_t* cur = img16;
        if (f > 0 && a->warp && a->warp[c]) {
          orc_warp_affine_u16(img16, w, h, a->warp[c] + (size_t)off * 6, a->interp, warp16, w, h);
          cur = warp16;
        }
        const int patched = a->n_clusters && a->n_clusters[c] > 0;
        if (a->filter_kind && !patched) {   /* the image is still CV_16U: OpenCV's integer filter */
          uint16_t* other = (cur == img16) ? warp16 : img16;
          orc_filter_u16(cur, other, w, h, a->filter_kind, a->filter_size);
          cur = other;
        }
        for (size_t i = 0; i < P; ++i) img32[i] = (float)cur[i];
        if (patched) {
          orc_patch_apply(img32, w, a->n_clusters[c], a->bounds_off[c], a->bx[c], a->by[c],
                          a->internal_off[c], a->ix[c], a->iy[c]);
          if (a->filter_kind) {
            float* t32 = (float*)malloc(sizeof(float) * P);
            orc_filter_f32(img32, t32, w, h, a->filter_kind, a->filter_size);
            memcpy(img32, t32, sizeof(float) * P);
            free(t32);
          }
        }
        orc_project_frame(a->rowptr[c], a->col[c], a->val[c], N, img32, csol);
        if (c == 0)
          memcpy(sol, csol, sizeof(float) * (size_t)N);
        else
          for (int i = 0; i < N; ++i) sol[i] = sol[i] + csol[i];
      }
      for (int i = 0; i < a->n_skipped; ++i) sol[a->skipped[i]] = NAN;
      for (int i = 0; i < N; ++i) {
        lsq[i] += (double)(sol[i] * sol[i]);
        lsum[i] += (double)sol[i];
      }
      if (a->remap) {
        memcpy(tmpsol, sol, sizeof(float) * (size_t)N);
        for (int i = 0; i < N; ++i) sol[i] = tmpsol[a->remap[i]];
      }
      memcpy(intensity + (size_t)off * N, sol, sizeof(float) * (size_t)N);
    }
#pragma omp critical
    for (int i = 0; i < N; ++i) {
      sumsq[i] += lsq[i];
      sum[i] += lsum[i];
    }
    free(lsum);
    free(lsq);
    free(sol);
    free(csol);
    free(tmpsol);
    free(img16);
    free(warp16);
    free(img32);
  }
}

/* a10: finals cpp/exec/psp_process.cpp:1930-1940 */
ORC_API void orc_phase1_finals(const double* sum, const double* sumsq, int n_nodes,
                               unsigned n_frames_total, const int* remap, float* avg,
                               float* rms) {
  float* t = remap ? (float*)malloc(sizeof(float) * (size_t)n_nodes) : NULL;
  for (int i = 0; i < n_nodes; ++i) {
    avg[i] = (float)(sum[i] / n_frames_total);
    rms[i] = (float)sqrt(sumsq[i] / n_frames_total);
  }
  if (remap) {
    memcpy(t, rms, sizeof(float) * (size_t)n_nodes);
    for (int i = 0; i < n_nodes; ++i) rms[i] = t[remap[i]];
    memcpy(t, avg, sizeof(float) * (size_t)n_nodes);
    for (int i = 0; i < n_nodes; ++i) avg[i] = t[remap[i]];
    free(t);
  }
}

/* ------------------------------------------------------------------------ */
/* a11: local_transpose cpp/exec/psp_process.cpp:647-689 (100x100 tiles)     */
/*      global_transpose :707-771, with the MPI ranks simulated in-process:  */
/*      src[r] is rank r's [F_r][N] slice, dst[s] is rank s's [N_s][F] slice */
/* ------------------------------------------------------------------------ */
ORC_API void orc_local_transpose(const float* src, int x_extent, int y_extent, float* dst) {
  const int bs = 100;
  int fx = x_extent / bs, fy = y_extent / bs;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
  for (int yb = 0; yb < fy + 1; ++yb)
    for (int xb = 0; xb < fx + 1; ++xb) {
      int ys = yb * bs, ye = (yb < fy) ? bs : y_extent - fy * bs;
      int xs = xb * bs, xe = (xb < fx) ? bs : x_extent - fx * bs;
      for (int jj = 0; jj < ye; ++jj)
        for (int ii = 0; ii < xe; ++ii)
          dst[(size_t)(xs + ii) * y_extent + ys + jj] = src[(size_t)(ys + jj) * x_extent + xs + ii];
    }
}

ORC_API void orc_global_transpose(const float* const* src, float* const* dst, int n_ranks,
                                  int n_nodes, int n_frames) {
  int* fs = (int*)malloc(sizeof(int) * 4 * (size_t)n_ranks);
  int *fe = fs + n_ranks, *ns = fe + n_ranks, *ne = ns + n_ranks;
  orc_apportion(n_frames, n_ranks, fs, fe);
  orc_apportion(n_nodes, n_ranks, ns, ne);
  for (int r = 0; r < n_ranks; ++r) { /* sender */
    float* temp = (float*)malloc(sizeof(float) * (size_t)n_nodes * (fe[r] > 0 ? fe[r] : 1));
    orc_local_transpose(src[r], n_nodes, fe[r], temp);
    for (int s = 0; s < n_ranks; ++s) { /* receiver */
      const float* msg = temp + (size_t)ns[s] * fe[r];
      for (long no = 0; no < ne[s]; ++no)
        for (long fo = 0; fo < fe[r]; ++fo)
          dst[s][(size_t)no * n_frames + fs[r] + fo] = msg[(size_t)no * fe[r] + fo];
    }
    free(temp);
  }
  free(fs);
}

/* ------------------------------------------------------------------------ */
/* a14: TransPolyFitter<float> cpp/lib/filtering.ipp:13-26 (ctor), :48-76    */
/* ------------------------------------------------------------------------ */
ORC_API void orc_transpoly_build(unsigned n_frames, unsigned degree, float* A /*F x (deg+1) col-major*/) {
  for (unsigned f = 0; f < n_frames; ++f)
    for (unsigned c = 0; c <= degree; ++c)
      A[(size_t)c * n_frames + f] = (float)pow((double)((float)f / (float)n_frames), (double)c);
}

/* fit[F] = A * (ColPivHouseholderQR(A).solve(data)); scratch: F*(ncoef+1) floats */
ORC_API void orc_transpoly_eval_fit(const float* A, unsigned n_frames, unsigned ncoef,
                                    const float* data, float* fit, float* coef_out,
                                    float* scratch) {
  float* QR = scratch;
  float* c = scratch + (size_t)n_frames * ncoef;
  float hcoef[16], coef[16];
  int transp[16];
  memcpy(QR, A, sizeof(float) * (size_t)n_frames * ncoef);
  int nz = orc_colpiv_qr_f32(QR, (int)n_frames, (int)ncoef, hcoef, transp);
  orc_colpiv_qr_solve_f32(QR, (int)n_frames, (int)ncoef, hcoef, transp, nz, data, coef, c);
  for (unsigned f = 0; f < n_frames; ++f) {
    float s = 0.0f;
    for (unsigned k = 0; k < ncoef; ++k) s += A[(size_t)k * n_frames + f] * coef[k];
    fit[f] = s;
  }
  if (coef_out) memcpy(coef_out, coef, sizeof(float) * ncoef);
}

/* a13: PaintCalibration::get_gain cpp/lib/non_cv_upsp.cpp:66-68 */
ORC_API float orc_get_gain(const float cal[6], float T, float Pss) {
  float a = cal[0], b = cal[1], c = cal[2], d = cal[3], e = cal[4], f = cal[5];
  return a + b * T + c * T * T + (d + e * T + f * T * T) * Pss;
}

/* ------------------------------------------------------------------------ */
/* a12: phase-2 node loop cpp/exec/psp_process.cpp:2460-2498 for one rank's  */
/*      node slice.  itrans [n_local][F] node-major intensity; per-node      */
/*      vectors are indexed by the *local* node.  Skipped (coverage==0)      */
/*      nodes leave their pressure row untouched (the reference `continue`s  */
/*      before the store, :2466-2472) and get NaN stats.                      */
/*      exact_fit != 0 replaces the float QR detrend by a float64 least-      */
/*      squares fit on the same float matrix (used to bound the reference's   */
/*      own float noise in tests; not a reference behaviour).                 */
/* ------------------------------------------------------------------------ */
static void exact_fit_f64(const float* A, unsigned F, unsigned nc, const float* data,
                          float* fit) {
  /* modified Gram-Schmidt twice in long double on the float-valued columns */
  long double* Q = (long double*)malloc(sizeof(long double) * (size_t)F * nc);
  for (unsigned k = 0; k < nc; ++k) {
    long double* q = Q + (size_t)k * F;
    for (unsigned f = 0; f < F; ++f) q[f] = (long double)A[(size_t)k * F + f];
    for (int pass = 0; pass < 2; ++pass)
      for (unsigned j = 0; j < k; ++j) {
        const long double* p = Q + (size_t)j * F;
        long double d = 0;
        for (unsigned f = 0; f < F; ++f) d += p[f] * q[f];
        for (unsigned f = 0; f < F; ++f) q[f] -= d * p[f];
      }
    long double n = 0;
    for (unsigned f = 0; f < F; ++f) n += q[f] * q[f];
    n = sqrtl(n);
    for (unsigned f = 0; f < F; ++f) q[f] /= n;
  }
  long double* acc = (long double*)calloc(F, sizeof(long double));
  for (unsigned k = 0; k < nc; ++k) {
    const long double* q = Q + (size_t)k * F;
    long double d = 0;
    for (unsigned f = 0; f < F; ++f) d += q[f] * (long double)data[f];
    for (unsigned f = 0; f < F; ++f) acc[f] += d * q[f];
  }
  for (unsigned f = 0; f < F; ++f) fit[f] = (float)acc[f];
  free(acc);
  free(Q);
}

ORC_API void orc_phase2(int n_local, unsigned n_frames, const float* itrans,
                        const float* avg_final, const float* coverage, const float* steady,
                        const float* model_temp, const float cal[6], float qbar, float ps,
                        unsigned degree, int exact_fit, float* ptrans, double* rms,
                        double* avg, double* gain) {
  unsigned nc = degree + 1;
  float* A = (float*)malloc(sizeof(float) * (size_t)n_frames * nc);
  orc_transpoly_build(n_frames, degree, A);
#pragma omp parallel
  {
    float* node_sol = (float*)malloc(sizeof(float) * n_frames);
    float* fit = (float*)malloc(sizeof(float) * n_frames);
    float* scratch = (float*)malloc(sizeof(float) * (size_t)n_frames * (nc + 1));
#pragma omp for
    for (int i = 0; i < n_local; ++i) {
      if (coverage[i] == 0) {
        rms[i] = NAN;
        avg[i] = NAN;
        gain[i] = NAN;
        continue;
      }
      float Pss = qbar * steady[i] + ps;
      double local_gain = (double)orc_get_gain(cal, model_temp[i], Pss);
      for (unsigned f = 0; f < n_frames; ++f)
        node_sol[f] = avg_final[i] / itrans[(size_t)i * n_frames + f];
      if (exact_fit)
        exact_fit_f64(A, n_frames, nc, node_sol, fit);
      else
        orc_transpoly_eval_fit(A, n_frames, nc, node_sol, fit, NULL, scratch);
      double lr = 0.0, la = 0.0;
      for (unsigned f = 0; f < n_frames; ++f) {
        float pressure = (float)((node_sol[f] - fit[f]) * local_gain);
        node_sol[f] = (float)(pressure * 12.0 * 12.0 / qbar);
        ptrans[(size_t)i * n_frames + f] = node_sol[f];
        lr += (double)(node_sol[f] * node_sol[f]);
        la += (double)node_sol[f];
      }
      rms[i] = lr;
      avg[i] = la;
      gain[i] = local_gain;
    }
    free(node_sol);
    free(fit);
    free(scratch);
  }
  free(A);
}

/* finals cpp/exec/psp_process.cpp:2540-2547 */
ORC_API void orc_phase2_finals(const double* rms, const double* avg, const double* gain, int n,
                               unsigned n_frames, float* rms_f, float* avg_f, float* gain_f) {
  for (int i = 0; i < n; ++i) {
    avg_f[i] = (float)(avg[i] / n_frames);
    rms_f[i] = (float)sqrt(rms[i] / n_frames);
    gain_f[i] = (float)gain[i];
  }
}

/* bench.py: torchrun exports OMP_NUM_THREADS=1; the CPU arm asks for the host's cores explicitly */
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
