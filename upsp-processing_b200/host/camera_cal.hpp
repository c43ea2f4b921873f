// camera_cal.hpp -- the external camera calibration file of psp_process, host side.
// Mirrors upsp::read_json_camera_calibration (cpp/lib/CameraCal.cpp:18-54): JSON keys cameraMatrix
// (3x3), distCoeffs (the first 4 are used there; 5 or 8 are accepted here as calibrate_camera does,
// :67-79), rmat (3x3), tvec (3), imageSize (2); rvec = cv::Rodrigues(rmat).
// cv::Rodrigues(matrix -> vector) first re-orthonormalises the matrix with an SVD; this restatement
// assumes rmat is orthonormal to rounding (it is written by the calibration tools with 16 digits) and
// applies the same axis-angle formulas, which reproduces cv2.Rodrigues to ~1e-15 (tests pin 1e-12).
#pragma once
#include <array>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/upsp_gpu.h"

namespace upsp_b200 {

/* all numbers of the (possibly nested) array stored under "key" in a JSON text */
inline std::vector<double> json_numbers(const std::string& text, const std::string& key) {
  const size_t k = text.find("\"" + key + "\"");
  if (k == std::string::npos) throw std::invalid_argument("calibration file has no \"" + key + "\"");
  size_t p = text.find('[', k);
  if (p == std::string::npos) throw std::invalid_argument("\"" + key + "\" is not an array");
  std::vector<double> out;
  int depth = 0;
  for (; p < text.size(); ++p) {
    const char c = text[p];
    if (c == '[') ++depth;
    else if (c == ']') {
      if (--depth == 0) break;
    } else if (c == '-' || c == '+' || (c >= '0' && c <= '9') || c == '.') {
      char* end = nullptr;
      out.push_back(std::strtod(text.c_str() + p, &end));
      p = (size_t)(end - text.c_str()) - 1;
    }
  }
  if (depth != 0) throw std::invalid_argument("unterminated array under \"" + key + "\"");
  return out;
}

/* cv::Rodrigues(R -> rvec) for an orthonormal R (row-major 3x3) */
inline void rodrigues_from_matrix(const double R[9], double r[3]) {
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  const double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  const double theta = std::acos(c);
  if (s < 1e-5) {
    if (c > 0) {
      r[0] = r[1] = r[2] = 0;
    } else {
      double t;
      t = (R[0] + 1) * 0.5; rx = std::sqrt(std::max(t, 0.));
      t = (R[4] + 1) * 0.5; ry = std::sqrt(std::max(t, 0.)) * (R[1] < 0 ? -1. : 1.);
      t = (R[8] + 1) * 0.5; rz = std::sqrt(std::max(t, 0.)) * (R[2] < 0 ? -1. : 1.);
      if (std::fabs(rx) < std::fabs(ry) && std::fabs(rx) < std::fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
      const double scale = theta / std::sqrt(rx * rx + ry * ry + rz * rz);
      r[0] = rx * scale; r[1] = ry * scale; r[2] = rz * scale;
    }
  } else {
    const double vth = 1. / (2 * s) * theta;
    r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
  }
}

/* cv::Rodrigues(rvec -> R), the arithmetic of the library's own fill_setup_cam */
inline void rodrigues_to_matrix(const double r[3], double R[9]) {
  const double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (theta < 2.220446049250313e-16) {
    for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c, itheta = 1.0 / theta;
  const double x = r[0] * itheta, y = r[1] * itheta, z = r[2] * itheta;
  const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
  const double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  for (int k = 0; k < 9; ++k) R[k] = c * ((k % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[k] + s * rx[k];
}

/* CameraCal::get_cam_center (cpp/lib/CameraCal.cpp:193-204): -R^T t */
inline std::array<double, 3> get_cam_center(const upsp_camera_model& cam) {
  double R[9];
  rodrigues_to_matrix(cam.rvec, R);
  std::array<double, 3> c{};
  for (int i = 0; i < 3; ++i) c[(size_t)i] = -(R[0 + i] * cam.tvec[0] + R[3 + i] * cam.tvec[1] + R[6 + i] * cam.tvec[2]);
  return c;
}

inline upsp_camera_model read_json_camera_calibration(const std::string& cfg_file) {
  std::ifstream ifs(cfg_file);
  if (!ifs) throw std::invalid_argument("Cannot open camera calibration file '" + cfg_file + "'");
  std::stringstream ss;
  ss << ifs.rdbuf();
  const std::string text = ss.str();
  const auto K = json_numbers(text, "cameraMatrix"), d = json_numbers(text, "distCoeffs"), R = json_numbers(text, "rmat"),
             t = json_numbers(text, "tvec"), sz = json_numbers(text, "imageSize");
  if (K.size() != 9 || R.size() != 9 || t.size() != 3 || sz.size() != 2)
    throw std::invalid_argument("camera calibration file '" + cfg_file + "': cameraMatrix/rmat 3x3, tvec 3, imageSize 2 expected");
  if (d.size() != 4 && d.size() != 5 && d.size() != 8)
    throw std::invalid_argument("distCoeffs should be a row vector with 4,5,or 8 coefficients");      // CameraCal.cpp:75-77
  upsp_camera_model cam{};
  cam.fx = K[0]; cam.fy = K[4]; cam.cx = K[2]; cam.cy = K[5];
  for (size_t i = 0; i < d.size(); ++i) cam.dist[i] = d[i];
  rodrigues_from_matrix(R.data(), cam.rvec);
  for (int i = 0; i < 3; ++i) cam.tvec[i] = t[i];
  cam.width = (int)sz[0];
  cam.height = (int)sz[1];
  return cam;
}

}  // namespace upsp_b200
