// run_inputs.hpp -- the small per-run input files psp_process reads around the frame chain, host side
// (SURVEY 8b process contract; 8f rank 2/3).  No OpenCV / Eigen / Boost.
//   PaintCalibration          cpp/lib/non_cv_upsp.cpp:19-68        a..f of the gain polynomial (`key = value` lines)
//   read_tunnel_conditions    cpp/lib/non_cv_upsp.cpp:109-200      the .wtd / sds file ('#' header line + value line)
//   model_temperature         cpp/exec/psp_process.cpp:2287-2310   recovery-factor wall temperature or TCAVG
//   read_psp_target_file      cpp/utils/file_readers.ipp:206-255   *Targets / *Fiducials sections of a .tgts file
//   read_plot3d_scalar_function_file   cpp/lib/plot3d.cpp:12-101   steady-state Cp / model temperature
//   set_surface_normals       cpp/utils/file_readers.ipp:12-90     csv (nidx, x_norm, y_norm, z_norm) overriding node normals
//   read_active_comp_file     cpp/utils/file_readers.cpp:12-48     csv (component, active) -> nodes of inactive components
//   intensity_histc           cpp/lib/image_processing.ipp:10-49   first-frame histogram
//   find_peaks, first_min_threshold    cpp/utils/clustering.ipp:9-101   -> the patch boundary threshold
//   patch_threshold           cpp/exec/psp_process.cpp:2150-2155   edges[first_min_threshold(counts, 5)] + 5
// Compile with -ffp-contract=off: the float expressions are written in the reference's order.
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <unordered_map>
#include <stdexcept>
#include <string>
#include <vector>

namespace upsp_b200 {

namespace detail {
inline std::string strip_all_space(std::string s) {
  s.erase(std::remove_if(s.begin(), s.end(), [](unsigned char c) { return std::isspace(c); }), s.end());
  return s;
}
/* std::getline(ss, seg, delim) semantics: no empty piece after a trailing delimiter */
inline std::vector<std::string> split_char(const std::string& s, char delim) {
  std::vector<std::string> out;
  std::stringstream ss(s);
  for (std::string seg; std::getline(ss, seg, delim);) out.push_back(seg);
  return out;
}
inline std::vector<std::string> split_ws(const std::string& s) {
  std::vector<std::string> out;
  std::istringstream ss(s);
  for (std::string t; ss >> t;) out.push_back(t);
  return out;
}
}  // namespace detail

/* ---- paint calibration ---- */
struct PaintCalibration {
  float a = 0.f, b = 0.f, c = 0.f, d = 0.f, e = 0.f, f = 0.f;
  PaintCalibration() = default;
  explicit PaintCalibration(const std::string& filename) {
    std::ifstream ifs(filename);
    if (!ifs) throw std::invalid_argument("Cannot read paint calibration file");
    for (std::string buf; std::getline(ifs, buf);) {
      const auto tokens = detail::split_char(detail::strip_all_space(buf), '=');
      if (tokens.size() != 2) continue;
      float var = 0;
      try {
        var = (float)std::stod(tokens[1]);
      } catch (...) {   // the reference aborts here; an exception reaches the same exit code 1 path of the driver
        throw std::invalid_argument("Error: Could not parse coefficient " + tokens[0] + ".  Expected float");
      }
      if (tokens[0] == "a") a = var;
      else if (tokens[0] == "b") b = var;
      else if (tokens[0] == "c") c = var;
      else if (tokens[0] == "d") d = var;
      else if (tokens[0] == "e") e = var;
      else if (tokens[0] == "f") f = var;
    }
  }
  float get_gain(float T, float Pss) const { return a + b * T + c * T * T + (d + e * T + f * T * T) * Pss; }
};

/* ---- tunnel conditions ---- */
struct TunnelConditions {
  static constexpr float nan() { return std::numeric_limits<float>::quiet_NaN(); }
  float alpha = nan(), beta = nan(), phi = nan(), mach = nan(), rey = nan(), ptot = nan(), qbar = nan(), ttot = nan(),
        ps = nan(), tcavg = nan();
  std::string test_id;
  int run = 0, seq = 0;
};

inline TunnelConditions read_tunnel_conditions(const std::string& filename, std::ostream* warn = &std::cout) {
  std::ifstream ifs(filename);
  if (!ifs) throw std::invalid_argument("Cannot open '" + filename + "'");
  TunnelConditions cond;
  for (std::string line; std::getline(ifs, line);) {
    const auto terms = detail::split_ws(line);
    if (terms.empty() || !(terms[0].size() == 1 && terms[0][0] == '#')) continue;
    std::getline(ifs, line);
    const auto vals = detail::split_ws(line);
    if (vals.size() != terms.size() - 1) throw std::invalid_argument("Failed to parse '" + filename + "'");
    for (size_t i = 1; i < terms.size(); ++i) {
      float* dst = nullptr;
      if (terms[i] == "ALPHA") dst = &cond.alpha;
      else if (terms[i] == "BETA") dst = &cond.beta;
      else if (terms[i] == "PHI") dst = &cond.phi;
      else if (terms[i] == "MACH") dst = &cond.mach;
      else if (terms[i] == "RNU") dst = &cond.rey;
      else if (terms[i] == "PTOT") dst = &cond.ptot;
      else if (terms[i] == "Q") dst = &cond.qbar;
      else if (terms[i] == "TTF") dst = &cond.ttot;
      else if (terms[i] == "PS") dst = &cond.ps;
      else if (terms[i] == "TCAVG") dst = &cond.tcavg;
      if (!dst) continue;
      try {
        *dst = (float)std::stod(vals[i - 1]);
      } catch (...) {
        std::cerr << "Unable to parse variable " << terms[i] << " in '" << filename << "'" << std::endl;
      }
    }
    break;
  }
  if (warn) {
    const std::pair<const char*, float> chk[] = {{"ALPHA", cond.alpha}, {"BETA", cond.beta}, {"PHI", cond.phi},
                                                 {"MACH", cond.mach},   {"REYN", cond.rey},  {"PTOT", cond.ptot},
                                                 {"Q", cond.qbar},      {"TTOT", cond.ttot}, {"PS", cond.ps}};
    for (const auto& c : chk)
      if (std::isnan(c.second)) *warn << "Warning: " << c.first << " was not read from '" << filename << "'" << std::endl;
  }
  return cond;
}

/* model temperature of phase 2 in deg F; `wall_temp` receives the recovery-factor estimate */
inline float model_temperature(const TunnelConditions& tc, float* wall_temp_out = nullptr) {
  const float r = 0.896f, gamma = 1.4f, F_to_R = 459.67f;   // Phase2Settings, psp_process.cpp:1092-1098
  float ttot = tc.ttot;
  ttot += F_to_R;
  float t_inf = (float)(ttot / (1.0 + (gamma - 1.0) * 0.5 * tc.mach * tc.mach));
  ttot -= F_to_R;
  t_inf -= F_to_R;
  const float wall_temp = r * (ttot - t_inf) + t_inf;
  if (wall_temp_out) *wall_temp_out = wall_temp;
  return std::isnan(tc.tcavg) ? wall_temp : tc.tcavg;
}

/* ---- targets ---- */
struct ModelTarget {
  int num = 0;
  double x = 0, y = 0, z = 0, diameter = 0;
};

inline bool read_psp_target_file(const std::string& target_file, std::vector<ModelTarget>& targs, bool planar = false,
                                 const std::string& label = "*Targets") {
  std::ifstream ifs(target_file);
  if (!ifs.is_open()) return false;
  targs.clear();
  double tmp = 0, x = 0, y = 0, z = 0, diam = 0;   // as in the reference: a short line keeps the previous line's values
  for (std::string line; std::getline(ifs, line);) {
    if (line.find(label) == std::string::npos) continue;
    while (std::getline(ifs, line)) {
      if (!line.empty() && line[0] == '*') break;
      std::stringstream ss(line);
      int targ_id = 0;
      ss >> targ_id;
      ss >> x >> y >> z >> tmp >> tmp >> tmp >> diam;
      ModelTarget t;
      t.num = targ_id;
      t.x = x;
      t.y = y;
      t.z = planar ? 0.0 : z;
      t.diameter = diam;
      targs.push_back(t);
    }
    break;
  }
  return true;
}

/* ---- surface normal overrides (structured models only; psp_process refuses the file for a TriModel) ---- */
inline int set_surface_normals(const std::string& normal_file, std::vector<float>& normals /* [N][3] */) {
  std::ifstream ifs(normal_file);
  if (!ifs) throw std::invalid_argument("Cannot open surface normal csv file");
  std::string line;
  std::getline(ifs, line);
  const auto header = detail::split_char(detail::strip_all_space(line), ',');
  int col[4] = {-1, -1, -1, -1};
  const char* names[4] = {"nidx", "x_norm", "y_norm", "z_norm"};
  for (size_t i = 0; i < header.size(); ++i)
    for (int k = 0; k < 4; ++k)
      if (header[i] == names[k]) col[k] = (int)i;
  for (int k = 0; k < 4; ++k)
    if (col[k] == -1) throw std::invalid_argument(std::string("Could not parse ") + names[k] + " in normal csv file");
  int nset = 0;
  while (std::getline(ifs, line)) {
    const auto terms = detail::split_char(detail::strip_all_space(line), ',');
    long nidx;
    float n[3];
    try {
      nidx = std::stoi(terms.at((size_t)col[0]));
      for (int k = 0; k < 3; ++k) n[k] = (float)std::stod(terms.at((size_t)col[k + 1]));
    } catch (...) {
      throw std::invalid_argument("Cannot parse normal csv file'" + normal_file + "'");
    }
    if (nidx < 0 || (size_t)nidx * 3 + 2 >= normals.size()) throw std::invalid_argument("normal csv file: node index outside the model");
    for (int k = 0; k < 3; ++k) normals[(size_t)nidx * 3 + k] = n[k];
    ++nset;
  }
  return nset;
}

/* ---- active components ---- */
inline std::unordered_map<int, bool> read_active_comp_file(const std::string& comp_file) {
  std::ifstream ifs(comp_file);
  if (!ifs) throw std::invalid_argument("Cannot open active component csv file");
  std::unordered_map<int, bool> active_comps;
  std::string line;
  std::getline(ifs, line);   // header
  while (std::getline(ifs, line)) {
    const auto terms = detail::split_char(line, ',');
    try {
      const int comp = std::stoi(terms.at(0));
      active_comps[comp] = std::abs(std::stoi(terms.at(1))) != 0;
    } catch (...) {
      throw std::invalid_argument("Cannot parse active component csv file");
    }
  }
  return active_comps;
}

/* ---- plot3d scalar function file ---- */
namespace detail {
template <typename T>
bool read_raw(std::ifstream& ifs, T* dst, size_t nv) {
  ifs.read(reinterpret_cast<char*>(dst), (std::streamsize)(sizeof(T) * nv));
  return (bool)ifs;
}
template <typename T>
bool read_record(std::ifstream& ifs, T* dst, size_t nv, bool with_seps) {
  if (!with_seps) return read_raw(ifs, dst, nv);
  int32_t sep = 0;
  const int32_t expect = (int32_t)(nv * sizeof(T));
  if (!read_raw(ifs, &sep, 1) || sep != expect) return false;
  if (!read_raw(ifs, dst, nv)) return false;
  return read_raw(ifs, &sep, 1) && sep == expect;
}
inline std::string read_p3d_function(const std::string& filename, std::vector<float>& sol, bool seps) {
  sol.clear();
  std::ifstream ifs(filename, std::ifstream::binary);
  if (!ifs) return "Failed to open file";
  int32_t number_zones = 0;
  if (!read_record(ifs, &number_zones, 1, seps)) return "Failed to parse number of zones";
  if (number_zones < 0 || number_zones > (1 << 24)) return "Failed to parse number of zones";
  std::vector<int32_t> sizes((size_t)number_zones * 4);
  if (!read_record(ifs, sizes.data(), sizes.size(), seps))
    return "Failed to parse zone sizes (expected " + std::to_string(number_zones) + " zones)";
  long total = 0;
  for (int z = 0; z < number_zones; ++z) total += (long)sizes[z * 4] * sizes[z * 4 + 1] * sizes[z * 4 + 2];
  if (total < 0) return "Failed to parse zone sizes";
  {   // a header that promises more scalars than the file holds fails here instead of in the allocator
    const std::streampos here = ifs.tellg();
    ifs.seekg(0, std::ios::end);
    const std::streamoff left = ifs.tellg() - here;
    ifs.seekg(here);
    if ((std::streamoff)total * 4 > left)
      return "failed to read scalars (expected " + std::to_string(number_zones) + " zones, " + std::to_string(total) + " scalars)";
  }
  sol.resize((size_t)total);
  // plot3d.cpp:68 calls the separator-less overload for the data record whatever `seps` is: with FORTRAN
  // record markers the first value read is the leading marker and every scalar sits one slot late.  Kept.
  if (!read_raw(ifs, sol.data(), sol.size()))
    return "failed to read scalars (expected " + std::to_string(number_zones) + " zones, " + std::to_string(total) + " scalars)";
  return "";
}
}  // namespace detail

/* record_seps: +1 with FORTRAN record markers, 0 without, -1 try with then without (plot3d.cpp:79-101) */
inline std::vector<float> read_plot3d_scalar_function_file(const std::string& filename, int record_seps = -1) {
  std::vector<float> sol;
  std::string err;
  if (record_seps == 1 || record_seps == -1) {
    const std::string e = detail::read_p3d_function(filename, sol, true);
    if (e.empty()) return sol;
    err += "\nAssuming *has* FORTRAN record separators: " + e;
  }
  if (record_seps == 0 || record_seps == -1) {
    const std::string e = detail::read_p3d_function(filename, sol, false);
    if (e.empty()) return sol;
    err += "\nAssuming *no* FORTRAN record separators: " + e;
  }
  throw std::invalid_argument("Failed to parse Plot3D function file '" + filename + "':" + err + "\n");
}

/* ---- regression samples: at most `maxels` evenly strided values of a result vector (vv-*.dat) ----
 * upsp::fwrite cpp/utils/file_writers.cpp:9-31 and the identical lambda at psp_process.cpp:1984-2005 */
inline int write_regression_sample(const std::string& fname, const float* v, size_t n, int maxels) {
  if (n == 0) return -1;
  const size_t numels = maxels > 0 ? (size_t)maxels : n;
  const size_t step = n < numels ? 1 : n / numels;
  std::vector<float> outp;
  for (size_t jj = 0; outp.size() < numels && jj < n; jj += step) outp.push_back(v[jj]);
  FILE* fp = std::fopen(fname.c_str(), "wb");
  if (!fp) return -1;
  const int res = (int)std::fwrite(outp.data(), sizeof(float), outp.size(), fp);
  std::fclose(fp);
  return res;
}

/* ---- first-frame histogram -> boundary threshold of the patcher ---- */
inline void intensity_histc(const uint16_t* img, size_t n_px, std::vector<int>& edges, std::vector<int>& counts,
                            unsigned depth = 12, int bins = -1) {
  edges.clear();
  counts.clear();
  if (depth > 16) depth = 16;
  const unsigned max_value = 1u << depth;
  if (bins == -1) bins = (int)max_value;
  const uint16_t bin_sz = (uint16_t)std::ceil(max_value / (unsigned)bins);   // integer division first, as upstream
  counts.assign((size_t)bins, 0);
  for (size_t i = 0; i < n_px; ++i)
    if (img[i] < max_value) ++counts[(size_t)(img[i] / bin_sz)];
  edges.resize((size_t)bins + 1);
  for (size_t i = 0; i < edges.size(); ++i) edges[i] = (int)(i * bin_sz);
}

template <typename T>
void find_peaks(const std::vector<T>& data, std::vector<unsigned>& peaks, unsigned separation = 0) {
  peaks.clear();
  if (data.size() < 3) return;
  bool plateau = false;
  unsigned plateau_begin = 0;
  for (unsigned i = 1; i < data.size() - 1; ++i) {
    if (std::isinf((double)data[i]) || (data[i] > data[i - 1] && data[i] > data[i + 1])) {
      if (!peaks.empty() && (i - peaks.back()) < separation) {
        if (data[peaks.back()] < data[i]) peaks.back() = i;
        break;   // upstream leaves the whole scan here
      }
      peaks.push_back(i);
    } else if (data[i] > data[i - 1] && data[i] == data[i + 1]) {
      plateau = true;
      plateau_begin = i;
    } else if (plateau) {
      if (data[i] < data[i + 1]) plateau = false;
      else if (data[i] > data[i + 1]) {
        plateau = false;
        const unsigned plateau_i = (i + plateau_begin) / 2;
        if (!peaks.empty() && (plateau_i - peaks.back()) < separation) {
          if (data[peaks.back()] < data[plateau_i]) peaks.back() = plateau_i;
          break;
        }
        peaks.push_back(plateau_i);
      }
    }
  }
}

inline unsigned first_min_threshold(const std::vector<int>& counts, unsigned separation = 1) {
  std::vector<unsigned> max_peaks, min_peaks;
  find_peaks(counts, max_peaks, separation);
  if (max_peaks.empty()) return 0;
  std::vector<double> inverse_counts(counts.size());
  for (size_t i = 0; i < counts.size(); ++i) inverse_counts[i] = 1.0 / counts[i];
  find_peaks(inverse_counts, min_peaks, separation);
  for (unsigned m : min_peaks)
    if (m > max_peaks[0]) return m;
  return 0;
}

/* the threshold InitializeImagePatches hands to PatchClusters::threshold_bounds */
inline unsigned patch_threshold(const uint16_t* first_frame, size_t n_px, unsigned bit_depth) {
  std::vector<int> edges, counts;
  intensity_histc(first_frame, n_px, edges, counts, bit_depth, 256);
  return (unsigned)(edges[first_min_threshold(counts, 5)] + 5);
}

}  // namespace upsp_b200
