// targets.hpp -- fiducial targets of the patcher on the host: which targets a camera sees, where,
// and how large they appear.  A run has a few dozen targets, so this stays host C++; the per-node version of the same
// geometry (hundreds of thousands of rays) is the GPU operator upsp_op_create_projection (csrc/kernels_setup.cuh),
// whose formulas these restate with plain float / double operators (compile with -ffp-contract=off).
//   map_point_to_image     cpp/lib/CameraCal.ipp:225-239 = cv::projectPoints (double arithmetic, float in / out)
//   Triangle::intersect    cpp/raycast/pspRT.cpp:110-181 (watertight test); the BVH only prunes: the answer is the hit
//                          with the smallest t over all triangles, found here by testing every triangle
//   getTargets             cpp/exec/psp_process.cpp:56-114
//   get_target_diameters   cpp/exec/psp_process.cpp:116-165, get_perpendicular cpp/utils/cv_extras.ipp:30-66
//   the call sequence of InitializeImagePatches  cpp/exec/psp_process.cpp:2095-2123  -> visible_targets()
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "camera_cal.hpp"
#include "patch_geometry.hpp"
#include "run_inputs.hpp"

namespace upsp_b200 {

struct HostCamera {
  double R[9], t[3], fx, fy, cx, cy, k[8];
  float orig[3];
  int width, height;
  explicit HostCamera(const upsp_camera_model& cam) {
    rodrigues_to_matrix(cam.rvec, R);
    for (int i = 0; i < 3; ++i) t[i] = cam.tvec[i];
    fx = cam.fx; fy = cam.fy; cx = cam.cx; cy = cam.cy;
    for (int i = 0; i < 8; ++i) k[i] = cam.dist[i];
    const auto c = get_cam_center(cam);
    for (int i = 0; i < 3; ++i) orig[i] = (float)c[(size_t)i];
    width = cam.width;
    height = cam.height;
  }
  /* cv::projectPoints for one point */
  void map_point_to_image(const float p[3], float& u, float& v) const {
    const double X = p[0], Y = p[1], Z = p[2];
    const double x0 = R[0] * X + R[1] * Y + R[2] * Z + t[0], y0 = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z != 0.0 ? 1.0 / z : 1.0;
    const double x = x0 * z, y = y0 * z;
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2.0 * x * y, a2 = r2 + 2.0 * x * x, a3 = r2 + 2.0 * y * y;
    const double cdist = 1.0 + k[0] * r2 + k[1] * r4 + k[4] * r6;
    const double icdist2 = 1.0 / (1.0 + k[5] * r2 + k[6] * r4 + k[7] * r6);
    const double xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2, yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1;
    u = (float)(xd * fx + cx);
    v = (float)(yd * fy + cy);
  }
};

namespace rt_detail {
inline float v3_length(const float d[3]) {   // Imath::Vec3<float>::length
  const float len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  if (len2 < 2.f * std::numeric_limits<float>::min()) {
    const float ax = std::fabs(d[0]), ay = std::fabs(d[1]), az = std::fabs(d[2]);
    float m = ax > ay ? ax : ay;
    if (az > m) m = az;
    if (m == 0.f) return 0.f;
    const float x = ax / m, y = ay / m, z = az / m;
    return m * std::sqrt(x * x + y * y + z * z);
  }
  return std::sqrt(len2);
}
struct Ray {
  float o[3], d[3], Sx, Sy, Sz;
  int kx, ky, kz;
  Ray(const float org[3], const float dir[3]) {
    for (int i = 0; i < 3; ++i) o[i] = org[i], d[i] = dir[i];
    const float ax = std::fabs(d[0]), ay = std::fabs(d[1]), az = std::fabs(d[2]);
    kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    kx = (kz + 1) % 3;
    ky = (kx + 1) % 3;
    if (d[kz] < 0.f) std::swap(kx, ky);
    Sx = d[kx] / d[kz];
    Sy = d[ky] / d[kz];
    Sz = 1.f / d[kz];
  }
};
inline bool tri_hit(const Ray& ray, const float* pa, const float* pb, const float* pc, float& t_out) {
  float A[3], B[3], C[3];
  for (int i = 0; i < 3; ++i) A[i] = pa[i] - ray.o[i], B[i] = pb[i] - ray.o[i], C[i] = pc[i] - ray.o[i];
  const float Ax = A[ray.kx] - ray.Sx * A[ray.kz], Ay = A[ray.ky] - ray.Sy * A[ray.kz];
  const float Bx = B[ray.kx] - ray.Sx * B[ray.kz], By = B[ray.ky] - ray.Sy * B[ray.kz];
  const float Cx = C[ray.kx] - ray.Sx * C[ray.kz], Cy = C[ray.ky] - ray.Sy * C[ray.kz];
  float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
  if (U == 0.f || V == 0.f || W == 0.f) {
    U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
    V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
    W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
  }
  if ((U < 0.f || V < 0.f || W < 0.f) && (U > 0.f || V > 0.f || W > 0.f)) return false;
  const float det = U + V + W;
  if (det == 0.f) return false;
  const float Az = ray.Sz * A[ray.kz], Bz = ray.Sz * B[ray.kz], Cz = ray.Sz * C[ray.kz];
  const float T = U * Az + V * Bz + W * Cz;
  float xorf_T = std::fabs(T);
  if (std::signbit(T) != std::signbit(det)) xorf_T = -xorf_T;
  const float abs_det = std::fabs(det);
  if (xorf_T < 0.0f * abs_det || std::numeric_limits<float>::infinity() * abs_det < xorf_T) return false;
  t_out = T * (1.f / det);
  return true;
}
/* nearest hit over all triangles; ties in t go to the lowest triangle id, as in the GPU operator */
inline bool intersect(const Ray& ray, const float* verts, const int32_t* tris, int n_tris, float& t_hit) {
  bool any = false;
  float best = std::numeric_limits<float>::max();
  for (int k = 0; k < n_tris; ++k) {
    float t;
    if (tri_hit(ray, verts + 3 * tris[3 * k], verts + 3 * tris[3 * k + 1], verts + 3 * tris[3 * k + 2], t) && (!any || t < best)) {
      any = true;
      best = t;
    }
  }
  t_hit = best;
  return any;
}
/* kd_nearest over every model node (double squared distance); equidistant nodes: the lowest index */
inline int nearest_node(const float* xyz, int n_nodes, const double p[3]) {
  int best = 0;
  double bd = std::numeric_limits<double>::infinity();
  for (int n = 0; n < n_nodes; ++n) {
    const double dx = xyz[3 * n] - p[0], dy = xyz[3 * n + 1] - p[1], dz = xyz[3 * n + 2] - p[2];
    const double d = dx * dx + dy * dy + dz * dz;
    if (d < bd) bd = d, best = n;
  }
  return best;
}
}  // namespace rt_detail

struct ImageTarget {
  int num = 0;
  float xyz[3] = {0, 0, 0};
  float u = 0, v = 0, diameter = 0;
};

/* targets that are in the frame, not hidden by the model and not seen too obliquely */
inline std::vector<ImageTarget> get_targets(const HostCamera& cam, const float* xyz, const float* normals, int n_nodes,
                                            const int32_t* tris, int n_tris, const std::vector<ImageTarget>& orig_targs,
                                            float oblique_thresh) {
  std::vector<ImageTarget> out;
  for (const ImageTarget& targ : orig_targs) {
    float px, py;
    cam.map_point_to_image(targ.xyz, px, py);
    if (px < 0 || py < 0 || px >= (float)cam.width || py >= (float)cam.height) continue;
    float dir[3] = {targ.xyz[0] - cam.orig[0], targ.xyz[1] - cam.orig[1], targ.xyz[2] - cam.orig[2]};
    const float dist_from_eye = rt_detail::v3_length(dir);
    if (dist_from_eye != 0.f)
      for (float& c : dir) c /= dist_from_eye;
    const rt_detail::Ray ray(cam.orig, dir);
    float t;
    if (!rt_detail::intersect(ray, xyz, tris, n_tris, t)) continue;
    if ((double)t < (double)dist_from_eye - 1e-3) continue;   // something lies between the camera and the target
    const double hitpos[3] = {(double)(ray.o[0] + t * ray.d[0]), (double)(ray.o[1] + t * ray.d[1]), (double)(ray.o[2] + t * ray.d[2])};
    const int nn = rt_detail::nearest_node(xyz, n_nodes, hitpos);
    const float cos_theta = normals[3 * nn] * dir[0] + normals[3 * nn + 1] * dir[1] + normals[3 * nn + 2] * dir[2];
    const float ang = (float)std::acos((double)cos_theta);
    if (ang > oblique_thresh) out.push_back(targ);
  }
  return out;
}

namespace rt_detail {
inline float cvnorm3(const float v[3]) { return (float)std::sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]); }
inline void get_perpendicular(const float vec_in[3], float out[3]) {
  out[0] = out[1] = out[2] = 0.f;
  const float norm = cvnorm3(vec_in);
  if (norm == 0.f) return;
  const float vec[3] = {vec_in[0] / norm, vec_in[1] / norm, vec_in[2] / norm};
  const float a0 = std::fabs(vec[0]), a1 = std::fabs(vec[1]), a2 = std::fabs(vec[2]);
  const int max_comp = a0 > a1 ? (a0 > a2 ? 0 : 2) : (a1 > a2 ? 1 : 2);
  if (max_comp == 0) {
    out[1] = 1.f;
    out[0] = -(out[1] * vec[1] + out[2] * vec[2]) / vec[0];
  } else if (max_comp == 1) {
    out[0] = 1.f;
    out[1] = -(out[0] * vec[0] + out[2] * vec[2]) / vec[1];
  } else {
    out[0] = 1.f;
    out[2] = -(out[0] * vec[0] + out[1] * vec[1]) / vec[2];
  }
  const double n = std::sqrt((double)out[0] * out[0] + (double)out[1] * out[1] + (double)out[2] * out[2]);
  for (int i = 0; i < 3; ++i) out[i] = (float)((double)out[i] / n);   // Point3f / double
}
}  // namespace rt_detail

/* apparent diameter in pixels: a circle of the target's size in the surface's tangent plane, 4 points projected */
inline std::vector<float> get_target_diameters(const HostCamera& cam, const float* xyz, const float* normals, int n_nodes,
                                               const std::vector<ImageTarget>& targs) {
  std::vector<float> diams(targs.size(), 0.f);
  for (size_t i = 0; i < targs.size(); ++i) {
    const ImageTarget& tg = targs[i];
    const long iu = std::lrintf(tg.u), iv = std::lrintf(tg.v);   // Point2f -> Point2i
    if (tg.diameter == 0.0f || !(iu >= 0 && iv >= 0 && iu < cam.width && iv < cam.height)) continue;
    const double pos[3] = {tg.xyz[0], tg.xyz[1], tg.xyz[2]};
    const float* normal = normals + 3 * (size_t)rt_detail::nearest_node(xyz, n_nodes, pos);
    float a[3], b[3];
    rt_detail::get_perpendicular(normal, a);
    b[0] = a[1] * normal[2] - a[2] * normal[1];
    b[1] = a[2] * normal[0] - a[0] * normal[2];
    b[2] = a[0] * normal[1] - a[1] * normal[0];
    float theta = 0.0f, out_diameter = 0.0f;
    for (int j = 0; j < 4; ++j) {
      const double ca = 0.5 * tg.diameter * std::cos(theta), sb = 0.5 * tg.diameter * std::sin(theta);   // cosf / sinf, widened
      float est[3];
      for (int d = 0; d < 3; ++d) est[d] = (tg.xyz[d] + (float)(ca * a[d])) + (float)(sb * b[d]);
      float pu, pv;
      cam.map_point_to_image(est, pu, pv);
      const float du = pu - tg.u, dv = pv - tg.v;
      out_diameter = (float)(out_diameter + 2.0 * std::sqrt((double)du * du + (double)dv * dv));
      theta = (float)(theta + 2 * 3.141592653589793 / 4);
    }
    diams[i] = (float)(out_diameter / 4.0);
  }
  return diams;
}

/* InitializeImagePatches up to the clustering: *Targets + *Fiducials of the file -> visible, projected, sized */
/* normals: Model::get_n() (getTargets); node_normals: Node::get_normal() (get_target_diameters) */
inline std::vector<Target> visible_targets(const HostCamera& cam, const float* xyz, const float* normals, const float* node_normals, int n_nodes,
                                           const int32_t* tris, int n_tris, const std::string& target_file, float oblique_angle,
                                           float target_diam_sf) {
  std::vector<ModelTarget> orig, fiducials;
  if (!read_psp_target_file(target_file, orig)) throw std::invalid_argument("Cannot open '" + target_file + "'");
  if (!read_psp_target_file(target_file, fiducials, false, "*Fiducials")) throw std::invalid_argument("Cannot open '" + target_file + "'");
  orig.insert(orig.end(), fiducials.begin(), fiducials.end());
  std::vector<ImageTarget> all;
  for (const ModelTarget& m : orig) {
    ImageTarget t;
    t.num = m.num;
    t.xyz[0] = (float)m.x;
    t.xyz[1] = (float)m.y;
    t.xyz[2] = (float)m.z;
    t.diameter = (float)m.diameter;
    all.push_back(t);
  }
  const float thresh = (float)((180. - std::min(oblique_angle + 5.0, 90.0)) * 3.141592653589793 / 180.0);
  std::vector<ImageTarget> vis = get_targets(cam, xyz, normals, n_nodes, tris, n_tris, all, thresh);
  for (ImageTarget& t : vis) cam.map_point_to_image(t.xyz, t.u, t.v);
  const std::vector<float> diams = get_target_diameters(cam, xyz, node_normals, n_nodes, vis);
  std::vector<Target> out;
  for (size_t i = 0; i < vis.size(); ++i) {
    Target t;
    t.u = vis[i].u;
    t.v = vis[i].v;
    t.diameter = diams[i] * target_diam_sf;
    out.push_back(t);
  }
  return out;
}

}  // namespace upsp_b200
