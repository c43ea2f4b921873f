// kernels_setup.cuh -- phase-0 product on the GPU: the pixel-to-node projection matrix
// (SURVEY 8f rank 2).  Reference: create_projection_mat cpp/exec/psp_process.cpp:168-355
// (cv::projectPoints via CameraCal::map_point_to_image cpp/lib/CameraCal.ipp:227-239, nearest-hit
// ray cast through rt::BVH cpp/raycast/pspRT.cpp:359-430 with the watertight triangle test :110-181,
// six jittered retries, obliqueness test, nearest pixel).
//
// One thread per model node.  The reference prunes the ray cast with a SAH BVH; every ray here
// starts at the camera centre, so the pruning structure is a 2-D grid over ray DIRECTIONS
// (pinhole coordinates x/z, y/z in the camera frame): a triangle is listed in every cell its
// projected bounding box (plus a margin far larger than the float error of the edge tests)
// overlaps, and a ray only tests the triangles of its own cell.  Triangles that reach behind the
// camera plane are listed separately and tested by every ray; rays pointing behind the camera
// plane test all triangles.  The result of a cast is the hit with the strictly smallest t, i.e.
// exactly what the reference's full BVH traversal returns.
//
// All float arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction), the double
// projection likewise, so a node's pixel and (u, v) match a CPU evaluation of the same formulas bit
// for bit.
#pragma once
#include "common.cuh"

namespace upsp {

struct SetupCam {
  double R[9], t[3];          // cv::Rodrigues(rvec), tvec
  double fx, fy, cx, cy;
  double k[8];                // k1 k2 p1 p2 k3 k4 k5 k6
  float orig[3];              // camera centre (-R^T t narrowed to float, CameraCal.cpp:193-204)
  int width, height;
};

struct SetupGrid {
  double x0, y0, inv_cell_x, inv_cell_y;   // direction-space grid: cell = floor((xn - x0) * inv_cell)
  int gx, gy;
  const int* cell_start;      // [gx*gy + 1]
  const int* cell_tris;       // triangle ids, cell after cell
  const int* always;          // triangles every ray must test (reach behind the camera plane)
  int n_always;
};

struct SetupRay {
  float o[3], d[3];
  int kx, ky, kz;
  float Sx, Sy, Sz;
};

__device__ __forceinline__ float sel3(const float (&v)[3], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }

__device__ __forceinline__ void setup_ray_init(SetupRay& r, const float (&o)[3], const float (&d)[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    r.o[i] = o[i];
    r.d[i] = d[i];
  }
  const float ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]);
  r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
  r.kx = r.kz + 1;
  if (r.kx == 3) r.kx = 0;
  r.ky = r.kx + 1;
  if (r.ky == 3) r.ky = 0;
  if (sel3(d, r.kz) < 0.f) {
    const int t = r.kx;
    r.kx = r.ky;
    r.ky = t;
  }
  r.Sx = __fdiv_rn(sel3(d, r.kx), sel3(d, r.kz));
  r.Sy = __fdiv_rn(sel3(d, r.ky), sel3(d, r.kz));
  r.Sz = __fdiv_rn(1.f, sel3(d, r.kz));
}

// rt::Triangle::intersect (pspRT.cpp:110-181)
__device__ __forceinline__ bool setup_tri_hit(const SetupRay& ray, const float* __restrict__ pa,
                                              const float* __restrict__ pb, const float* __restrict__ pc, float& t_out) {
  float A[3], B[3], C[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    A[i] = __fsub_rn(__ldg(pa + i), ray.o[i]);
    B[i] = __fsub_rn(__ldg(pb + i), ray.o[i]);
    C[i] = __fsub_rn(__ldg(pc + i), ray.o[i]);
  }
  const float Akz = sel3(A, ray.kz), Bkz = sel3(B, ray.kz), Ckz = sel3(C, ray.kz);
  const float Ax = __fsub_rn(sel3(A, ray.kx), __fmul_rn(ray.Sx, Akz)), Ay = __fsub_rn(sel3(A, ray.ky), __fmul_rn(ray.Sy, Akz));
  const float Bx = __fsub_rn(sel3(B, ray.kx), __fmul_rn(ray.Sx, Bkz)), By = __fsub_rn(sel3(B, ray.ky), __fmul_rn(ray.Sy, Bkz));
  const float Cx = __fsub_rn(sel3(C, ray.kx), __fmul_rn(ray.Sx, Ckz)), Cy = __fsub_rn(sel3(C, ray.ky), __fmul_rn(ray.Sy, Ckz));
  float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
  float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
  float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
  if (U == 0.f || V == 0.f || W == 0.f) {
    U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
    V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
    W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
  }
  if ((U < 0.f || V < 0.f || W < 0.f) && (U > 0.f || V > 0.f || W > 0.f)) return false;
  const float det = __fadd_rn(__fadd_rn(U, V), W);
  if (det == 0.f) return false;
  const float Az = __fmul_rn(ray.Sz, Akz), Bz = __fmul_rn(ray.Sz, Bkz), Cz = __fmul_rn(ray.Sz, Ckz);
  const float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));
  float xorf_T = fabsf(T);
  if (signbit(T) != signbit(det)) xorf_T = -xorf_T;
  const float abs_det = fabsf(det);
  // ray_near = 0, hit_t = inf: (xorf_T < 0 * abs_det) or (inf * abs_det < xorf_T)
  if (xorf_T < __fmul_rn(0.0f, abs_det) || __fmul_rn(__int_as_float(0x7f800000), abs_det) < xorf_T) return false;
  t_out = __fmul_rn(T, __fdiv_rn(1.f, det));
  return true;
}

// nearest hit of one ray: primID, -2 = no hit
__device__ int setup_cast(const SetupRay& ray, const SetupCam& cam, const SetupGrid& g, const float* __restrict__ verts,
                          const int* __restrict__ tri, int n_tri) {
  float best = 3.402823466e+38f;
  int prim = -1;
  bool any = false;
  auto test = [&](int k) {
    float t;
    if (setup_tri_hit(ray, verts + 3 * __ldg(tri + 3 * k), verts + 3 * __ldg(tri + 3 * k + 1),
                      verts + 3 * __ldg(tri + 3 * k + 2), t)) {
      any = true;
      if (t < best || (t == best && k < prim)) {     // candidates arrive out of triangle order: ties go to the lowest id
        best = t;
        prim = k;
      }
    }
  };
  // direction in the camera frame (relative to the ray origin, which is the camera centre)
  const double dx = (double)ray.d[0], dy = (double)ray.d[1], dz = (double)ray.d[2];
  const double qx = cam.R[0] * dx + cam.R[1] * dy + cam.R[2] * dz;
  const double qy = cam.R[3] * dx + cam.R[4] * dy + cam.R[5] * dz;
  const double qz = cam.R[6] * dx + cam.R[7] * dy + cam.R[8] * dz;
  bool brute = !(qz > 0.0) || g.gx == 0;
  int cell = -1;
  if (!brute) {
    const double cxf = floor((qx / qz - g.x0) * g.inv_cell_x), cyf = floor((qy / qz - g.y0) * g.inv_cell_y);
    if (cxf >= 0.0 && cyf >= 0.0 && cxf < (double)g.gx && cyf < (double)g.gy) cell = (int)cyf * g.gx + (int)cxf;
    else brute = true;       // outside the gridded range (cannot happen for node rays; jittered rays at the rim)
  }
  if (brute) {
    for (int k = 0; k < n_tri; ++k) test(k);
  } else {
    for (int i = __ldg(g.cell_start + cell); i < __ldg(g.cell_start + cell + 1); ++i) test(__ldg(g.cell_tris + i));
    for (int i = 0; i < g.n_always; ++i) test(__ldg(g.always + i));
  }
  return any ? prim : -2;
}

__device__ __forceinline__ void setup_project(const SetupCam& c, const float* __restrict__ p, float& u, float& v) {
  const double X = (double)p[0], Y = (double)p[1], Z = (double)p[2];
  auto dot3 = [&](int r, int ti) {
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c.R[r], X), __dmul_rn(c.R[r + 1], Y)), __dmul_rn(c.R[r + 2], Z)), c.t[ti]);
  };
  const double x0 = dot3(0, 0), y0 = dot3(3, 1);
  double z = dot3(6, 2);
  z = z != 0.0 ? __ddiv_rn(1.0, z) : 1.0;
  const double x = __dmul_rn(x0, z), y = __dmul_rn(y0, z);
  const double* k = c.k;
  const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), r4 = __dmul_rn(r2, r2), r6 = __dmul_rn(r4, r2);
  const double a1 = __dmul_rn(__dmul_rn(2.0, x), y);
  const double a2 = __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)), a3 = __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y));
  const double cdist = __dadd_rn(__dadd_rn(__dadd_rn(1.0, __dmul_rn(k[0], r2)), __dmul_rn(k[1], r4)), __dmul_rn(k[4], r6));
  const double icdist2 = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dadd_rn(1.0, __dmul_rn(k[5], r2)), __dmul_rn(k[6], r4)), __dmul_rn(k[7], r6)));
  const double xd = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(x, cdist), icdist2), __dmul_rn(k[2], a1)), __dmul_rn(k[3], a2));
  const double yd = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(y, cdist), icdist2), __dmul_rn(k[2], a3)), __dmul_rn(k[3], a1));
  u = (float)__dadd_rn(__dmul_rn(xd, c.fx), c.cx);
  v = (float)__dadd_rn(__dmul_rn(yd, c.fy), c.cy);
}

__global__ void __launch_bounds__(128)
k_project_points(const SetupCam cam, const float* __restrict__ xyz, int n, float* __restrict__ uv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  setup_project(cam, xyz + 3 * i, uv[2 * i], uv[2 * i + 1]);
}

__global__ void __launch_bounds__(128)
k_create_projection(const SetupCam cam, const SetupGrid grid, const float* __restrict__ verts,
                    const float* __restrict__ normals, const uint8_t* __restrict__ is_data, int n_nodes,
                    const int* __restrict__ tri, int n_tri, float oblique_thresh, int* __restrict__ code,
                    float* __restrict__ uv) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  code[n] = -1;
  uv[2 * n] = 0.f;
  uv[2 * n + 1] = 0.f;
  if (!is_data[n]) return;
  const float ipos[3] = {__ldg(verts + 3 * n), __ldg(verts + 3 * n + 1), __ldg(verts + 3 * n + 2)};
  float ptx, pty;
  setup_project(cam, ipos, ptx, pty);
  if (isnan(ptx) || isnan(pty)) return;
  // upsp::contains(cv::Size, cv::Point2i(pt)): Point2f -> Point2i rounds half to even (cvRound)
  const int px = __float2int_rn(ptx), py = __float2int_rn(pty);
  if (!(px >= 0 && py >= 0 && px < cam.width && py < cam.height)) return;
  float dir[3] = {__fsub_rn(ipos[0], cam.orig[0]), __fsub_rn(ipos[1], cam.orig[1]), __fsub_rn(ipos[2], cam.orig[2])};
  {  // Imath::V3f::normalize
    const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(dir[0], dir[0]), __fmul_rn(dir[1], dir[1])), __fmul_rn(dir[2], dir[2]));
    float l = __fsqrt_rn(len2);
    if (len2 < 2.f * 1.175494351e-38f) {
      const float ax = fabsf(dir[0]), ay = fabsf(dir[1]), az = fabsf(dir[2]);
      float m = ax > ay ? ax : ay;
      if (az > m) m = az;
      if (m == 0.f) l = 0.f;
      else {
        const float x = __fdiv_rn(ax, m), y = __fdiv_rn(ay, m), z2 = __fdiv_rn(az, m);
        l = __fmul_rn(m, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z2, z2))));
      }
    }
    if (l != 0.f) {
      dir[0] = __fdiv_rn(dir[0], l);
      dir[1] = __fdiv_rn(dir[1], l);
      dir[2] = __fdiv_rn(dir[2], l);
    }
  }
  SetupRay ray;
  setup_ray_init(ray, cam.orig, dir);
  const int prim = setup_cast(ray, cam, grid, verts, tri, n_tri);
  if (prim == -2) return;
  auto has_node = [&](int p) { return __ldg(tri + 3 * p) == n || __ldg(tri + 3 * p + 1) == n || __ldg(tri + 3 * p + 2) == n; };
  bool visible = prim >= 0 && has_node(prim);
  if (!visible) {
    const float L = 1e-4f;
    for (int t = 0; !visible && t < 6; ++t) {
      const float s = (t & 1) ? L : -L;
      const float pos2[3] = {__fadd_rn(ipos[0], t < 2 ? s : 0.f), __fadd_rn(ipos[1], (t >> 1) == 1 ? s : 0.f),
                             __fadd_rn(ipos[2], t >= 4 ? s : 0.f)};
      const float dir2[3] = {__fsub_rn(pos2[0], cam.orig[0]), __fsub_rn(pos2[1], cam.orig[1]), __fsub_rn(pos2[2], cam.orig[2])};
      SetupRay ray2;
      setup_ray_init(ray2, cam.orig, dir2);      // un-normalised direction, as psp_process.cpp:273-275
      const int p2 = setup_cast(ray2, cam, grid, verts, tri, n_tri);
      if (p2 < 0) continue;
      visible = has_node(p2);
    }
  }
  if (!visible) return;
  const float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(normals + 3 * n), dir[0]), __fmul_rn(__ldg(normals + 3 * n + 1), dir[1])),
                                    __fmul_rn(__ldg(normals + 3 * n + 2), dir[2]));
  const float theta = (float)acos((double)cos_theta);    // acosf, correctly rounded
  if (!(theta > oblique_thresh)) return;
  uv[2 * n] = __fdiv_rn(ptx, (float)cam.width);
  uv[2 * n + 1] = __fdiv_rn(pty, (float)cam.height);
  const int rx = (int)roundf(ptx), ry = (int)roundf(pty);
  code[n] = ry * cam.width + rx;
}

}  // namespace upsp
