/* upsp_oracle_setup.c -- CPU restatement of psp_process's phase-0 projection-matrix construction
 * (TEST INFRASTRUCTURE ONLY: nothing under upsp-processing_b200/ may link or call it).
 *
 *   create_projection_mat      cpp/exec/psp_process.cpp:168-355
 *   CameraCal::map_point_to_image -> cv::projectPoints   cpp/lib/CameraCal.ipp:227-239
 *   CameraCal::get_cam_center  cpp/lib/CameraCal.cpp:193-204
 *   rt::Ray / rt::Triangle::intersect (watertight test)  cpp/raycast/pspRT.cpp:43-200
 *   rt::BVH::intersect         cpp/raycast/pspRT.cpp:359-430: a full traversal that keeps the hit with
 *       the strictly smallest t, i.e. the nearest hit over ALL triangles -- restated here as a
 *       brute-force loop in triangle order (the BVH only prunes; ties in t are broken by traversal
 *       order there and by triangle index here, which can only matter for coincident geometry).
 *
 * Third-party arithmetic: cv::projectPoints is OpenCV's (vcpkg opencv4); the published
 * pinhole + Brown-Conrady model of calib3d (cvProjectPoints2: x' = X/Z, radial 1+k1 r^2+k2 r^4+k3 r^6
 * over 1+k4 r^2+k5 r^4+k6 r^6, tangential p1, p2) is restated in double and pinned against
 * cv2.projectPoints 4.13 golden vectors (tests/golden/setup_golden.npz).  Imath::V3f::normalize is
 * restated as sqrt of the float dot product followed by three float divisions.
 * The ray casting itself is "parity unpinned": the reference has no test for it and its BVH cannot
 * be built here (Imath). */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct {
  double rvec[3], tvec[3];
  double fx, fy, cx, cy;
  double k[8];            /* k1 k2 p1 p2 k3 k4 k5 k6 (OpenCV order), unused ones 0 */
  int width, height;
} orc_camera;

/* cv::Rodrigues(rvec -> R), double (calib3d: theta = |r|, R = cos I + (1-cos) r r^T + sin [r]x) */
ORC_API void orc_rodrigues(const double r[3], double R[9]) {
  const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (theta < DBL_EPSILON) {
    memset(R, 0, 9 * sizeof(double));
    R[0] = R[4] = R[8] = 1.0;
    return;
  }
  const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, itheta = 1.0 / theta;
  const double x = r[0] * itheta, y = r[1] * itheta, z = r[2] * itheta;
  const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
  const double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int k = 0; k < 9; ++k) R[k] = c * I[k] + c1 * rrt[k] + s * rx[k];
}

/* cv::projectPoints for one float point; result as cv::Point2f */
static void project_point(const orc_camera* cam, const double R[9], const float p[3], float uv[2]) {
  const double X = p[0], Y = p[1], Z = p[2];
  const double x0 = R[0] * X + R[1] * Y + R[2] * Z + cam->tvec[0];
  const double y0 = R[3] * X + R[4] * Y + R[5] * Z + cam->tvec[1];
  double z = R[6] * X + R[7] * Y + R[8] * Z + cam->tvec[2];
  z = z ? 1.0 / z : 1.0;
  const double x = x0 * z, y = y0 * z;
  const double* k = cam->k;
  const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
  const double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
  const double cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6;
  const double icdist2 = 1. / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6);
  const double xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2;
  const double yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1;
  uv[0] = (float)(xd * cam->fx + cam->cx);
  uv[1] = (float)(yd * cam->fy + cam->cy);
}

ORC_API void orc_project_points(const orc_camera* cam, const float* xyz, int n, float* uv) {
  double R[9];
  orc_rodrigues(cam->rvec, R);
  for (int i = 0; i < n; ++i) project_point(cam, R, xyz + 3 * i, uv + 2 * i);
}

/* CameraCal::get_cam_center: -R^T t in double, narrowed to float by the caller's Point3_<float> */
ORC_API void orc_cam_center(const orc_camera* cam, float c[3]) {
  double R[9];
  orc_rodrigues(cam->rvec, R);
  for (int i = 0; i < 3; ++i)
    c[i] = (float)(-(R[0 + i] * cam->tvec[0] + R[3 + i] * cam->tvec[1] + R[6 + i] * cam->tvec[2]));
}

/* ---- rt::Ray (pspRT.cpp:45-71) ---- */
typedef struct {
  float o[3], d[3];
  int kx, ky, kz;
  float Sx, Sy, Sz;
} orc_ray;

static void ray_init(orc_ray* r, const float o[3], const float d[3]) {
  memcpy(r->o, o, sizeof r->o);
  memcpy(r->d, d, sizeof r->d);
  const float ad[3] = {fabsf(d[0]), fabsf(d[1]), fabsf(d[2])};
  r->kz = (ad[0] > ad[1]) ? (ad[0] > ad[2] ? 0 : 2) : (ad[1] > ad[2] ? 1 : 2);
  r->kx = r->kz + 1;
  if (r->kx == 3) r->kx = 0;
  r->ky = r->kx + 1;
  if (r->ky == 3) r->ky = 0;
  if (d[r->kz] < 0.f) {
    const int t = r->kx;
    r->kx = r->ky;
    r->ky = t;
  }
  r->Sx = d[r->kx] / d[r->kz];
  r->Sy = d[r->ky] / d[r->kz];
  r->Sz = 1.f / d[r->kz];
}

/* rt::Triangle::intersect (pspRT.cpp:110-181): hit distance t, or no hit */
static int tri_intersect(const orc_ray* ray, const float* pa, const float* pb, const float* pc, float* t_out) {
  float A[3], B[3], C[3];
  for (int i = 0; i < 3; ++i) {
    A[i] = pa[i] - ray->o[i];
    B[i] = pb[i] - ray->o[i];
    C[i] = pc[i] - ray->o[i];
  }
  const float Ax = A[ray->kx] - ray->Sx * A[ray->kz], Ay = A[ray->ky] - ray->Sy * A[ray->kz];
  const float Bx = B[ray->kx] - ray->Sx * B[ray->kz], By = B[ray->ky] - ray->Sy * B[ray->kz];
  const float Cx = C[ray->kx] - ray->Sx * C[ray->kz], Cy = C[ray->ky] - ray->Sy * C[ray->kz];
  float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
  if (U == 0.f || V == 0.f || W == 0.f) {
    U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
    V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
    W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
  }
  if ((U < 0.f || V < 0.f || W < 0.f) && (U > 0.f || V > 0.f || W > 0.f)) return 0;
  const float det = U + V + W;
  if (det == 0.f) return 0;
  const float Az = ray->Sz * A[ray->kz], Bz = ray->Sz * B[ray->kz], Cz = ray->Sz * C[ray->kz];
  const float T = U * Az + V * Bz + W * Cz;
  float xorf_T = fabsf(T);
  if (signbit(T) != signbit(det)) xorf_T = -xorf_T;
  const float abs_det = fabsf(det);
  if (xorf_T < 0.0f * abs_det || INFINITY * abs_det < xorf_T) return 0;
  const float rcpDet = 1.f / det;
  *t_out = T * rcpDet;
  return 1;
}

/* nearest hit over all triangles (rt::BVH::intersect semantics); returns primID or -1 */
static int nearest_hit(const orc_ray* ray, const float* verts, const int32_t* tri, int n_tri) {
  float best = FLT_MAX;
  int prim = -1, any = 0;
  for (int k = 0; k < n_tri; ++k) {
    float t;
    if (tri_intersect(ray, verts + 3 * tri[3 * k], verts + 3 * tri[3 * k + 1], verts + 3 * tri[3 * k + 2], &t)) {
      any = 1;
      if (t < best) {
        best = t;
        prim = k;
      }
    }
  }
  return any ? prim : -2;      /* -2: no hit at all; -1 cannot happen unless every hit had t >= FLT_MAX */
}

/* test hook: nearest hit of n rays (o[3], d[3] each) over a triangle soup; hit[i] = 0/1, t[i], prim[i] (lowest id on ties) */
ORC_API void orc_cast_rays(const float* verts, const int32_t* tri, int n_tri, const float* rays, int n, int32_t* hit, float* t_out,
                           int32_t* prim_out) {
  for (int i = 0; i < n; ++i) {
    orc_ray ray;
    ray_init(&ray, rays + 6 * i, rays + 6 * i + 3);
    float best = FLT_MAX;
    int prim = -1, any = 0;
    for (int k = 0; k < n_tri; ++k) {
      float t;
      if (tri_intersect(&ray, verts + 3 * tri[3 * k], verts + 3 * tri[3 * k + 1], verts + 3 * tri[3 * k + 2], &t)) {
        any = 1;
        if (t < best) best = t, prim = k;
      }
    }
    hit[i] = any;
    t_out[i] = best;
    prim_out[i] = prim;
  }
}

static void v3_normalize(float v[3]) {      /* Imath::Vec3<float>::normalize */
  const float len2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  float l = sqrtf(len2);
  if (len2 < 2.f * FLT_MIN) {               /* lengthTiny(): rescale by the largest component */
    const float ax = fabsf(v[0]), ay = fabsf(v[1]), az = fabsf(v[2]);
    float m = ax > ay ? ax : ay;
    if (az > m) m = az;
    if (m == 0.f) return;
    const float x = ax / m, y = ay / m, z2 = az / m;
    l = m * sqrtf(x * x + y * y + z2 * z2);
  }
  if (l != 0.f) {
    v[0] /= l;
    v[1] /= l;
    v[2] /= l;
  }
}

/* create_projection_mat: code[n] = pixel index (y*W + x) of node n, or -1; uv[2n] as the reference.
 * verts: [n_nodes][3] node positions (triangle vertices index nodes: triNodes == triangle corners),
 * normals: [n_nodes][3], is_data: [n_nodes] (Node::is_datanode), tri: [n_tri][3] node indices. */
ORC_API void orc_create_projection(const orc_camera* cam, const float* verts, const float* normals,
                                   const uint8_t* is_data, int n_nodes, const int32_t* tri, int n_tri,
                                   float oblique_thresh, int32_t* code, float* uv) {
  double R[9];
  orc_rodrigues(cam->rvec, R);
  float orig[3];
  orc_cam_center(cam, orig);
#pragma omp parallel for schedule(dynamic, 64)
  for (int n = 0; n < n_nodes; ++n) {
    code[n] = -1;
    uv[2 * n] = uv[2 * n + 1] = 0.f;
    if (!is_data[n]) continue;
    const float* ipos = verts + 3 * n;
    float pt[2];
    project_point(cam, R, ipos, pt);
    /* upsp::contains(cv::Size, cv::Point2i(pt)): Point2f -> Point2i is saturate_cast = cvRound */
    const long px = lrintf(pt[0]), py = lrintf(pt[1]);
    if (!(px >= 0 && py >= 0 && px < cam->width && py < cam->height)) continue;
    float dir[3] = {ipos[0] - orig[0], ipos[1] - orig[1], ipos[2] - orig[2]};
    v3_normalize(dir);
    orc_ray ray;
    ray_init(&ray, orig, dir);
    int prim = nearest_hit(&ray, verts, tri, n_tri);
    if (prim == -2) continue;
    int visible = prim >= 0 && (tri[3 * prim] == n || tri[3 * prim + 1] == n || tri[3 * prim + 2] == n);
    if (!visible) {
      const float L = 1e-4f;
      const float sp[6][3] = {{-L, 0, 0}, {L, 0, 0}, {0, -L, 0}, {0, L, 0}, {0, 0, -L}, {0, 0, L}};
      for (int t = 0; !visible && t < 6; ++t) {
        const float pos2[3] = {ipos[0] + sp[t][0], ipos[1] + sp[t][1], ipos[2] + sp[t][2]};
        const float dir2[3] = {pos2[0] - orig[0], pos2[1] - orig[1], pos2[2] - orig[2]};
        orc_ray ray2;
        ray_init(&ray2, orig, dir2);       /* built from the un-normalised direction (:273-275) */
        const int p2 = nearest_hit(&ray2, verts, tri, n_tri);
        if (p2 < 0) continue;
        visible = tri[3 * p2] == n || tri[3 * p2 + 1] == n || tri[3 * p2 + 2] == n;
      }
    }
    if (!visible) continue;
    const float* nn = normals + 3 * n;
    const float cos_theta = nn[0] * dir[0] + nn[1] * dir[1] + nn[2] * dir[2];
    const float theta = (float)acos((double)cos_theta);
    if (!(theta > oblique_thresh)) continue;
    uv[2 * n] = pt[0] / (float)cam->width;
    uv[2 * n + 1] = pt[1] / (float)cam->height;
    const int rx = (int)round((double)pt[0]), ry = (int)round((double)pt[1]);
    code[n] = ry * cam->width + rx;
  }
}
