// upsp_inputs.hpp -- the psp_process input deck (process contract), host side.
// Mirrors upsp::FileInputs (cpp/lib/upsp_inputs.cpp:35-173 Load, :343-700 block loaders, parse_line,
// evaluate_vars): '#' comments, '%Version' meta line, blocks @general @vars @all @camera @options
// @output of `key = value` lines (all whitespace is removed before splitting at '='), cameras sorted
// by number, $variables substituted (with the reference's own std::string::replace call, length
// argument included), @all targets / calibration filled into cameras that have none.
// Same defaults (upsp_inputs.cpp:27-33), same failure cases (Load returns false / throws).
#pragma once
#include <algorithm>
#include <cctype>
#include <fstream>
#include <numeric>
#include <cerrno>
#include <cstring>
#include <ostream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace upsp_b200 {

enum class TargetPatchType { None, Polynomial };
enum class RegistrationType { None, Point, Pixel };
enum class PixelInterpolationType { Linear, Nearest };
enum class FilterType { None, Gaussian, Box };
enum class OverlapKind { BestView, AverageViews };
enum class GridType { None, P3D, Tri };

struct FileInputs {
  std::string filename, version, test_id, tunnel, sds, grid, normals, grid_units = "-", active_comps, out_dir, out_name;
  int run = 0, sequence = 0;
  unsigned cameras = 0;
  std::vector<unsigned> cam_nums;
  std::vector<std::string> camera_filenames, cals, targets, vars, vars_map;
  GridType grid_type = GridType::None;
  TargetPatchType target_patcher = TargetPatchType::None;
  PixelInterpolationType pixel_interpolation = PixelInterpolationType::Linear;
  RegistrationType registration = RegistrationType::None;
  FilterType filter = FilterType::None;
  OverlapKind overlap = OverlapKind::AverageViews;
  int filter_size = 0, number_frames = 0;
  float oblique_angle = 70.f;
  std::string error;              // what LOG_ERROR would have printed when Load returns false

  bool Load(const std::string& input_file) {
    filename = input_file;
    std::ifstream fs(input_file);
    if (!fs.is_open()) return fail("Input file '" + input_file + "': cannot be opened");
    std::vector<std::string> lines;
    for (std::string l; std::getline(fs, l);) lines.push_back(l);
    bool fill_extras = false;
    std::string fill_targets, fill_calibration;
    size_t i = 0;
    // a block runs until the next line containing '@' (that line is re-read as a block header)
    auto block = [&](auto&& on_pair) {
      for (; i < lines.size(); ++i) {
        if (lines[i].find('@') != std::string::npos) return;
        std::vector<std::string> tok = parse_line(lines[i]);
        if (tok.size() == 2) on_pair(tok[0], tok[1]);
      }
    };
    bool ok = true;
    while (i < lines.size() && ok) {
      std::string buf = lines[i++];
      buf.erase(buf.begin(), std::find_if_not(buf.begin(), buf.end(), [](unsigned char c) { return std::isspace(c); }));
      if (buf.empty() || buf[0] == '#') continue;
      if (buf[0] == '%') {
        if (buf.find("Version") != std::string::npos) {
          std::vector<std::string> terms = split_whitespace(buf);
          if (terms.size() > 1) version = terms[1];
        }
        continue;
      }
      if (buf.find("@general") != std::string::npos) {
        block([&](const std::string& k, const std::string& v) {
          if (k == "test") test_id = v;
          else if (k == "run") ok = ok && to_int(v, run, "Error: Could not parse @general:run. Expected integer");
          else if (k == "sequence") ok = ok && to_int(v, sequence, "Error: Could not parse @general:sequence. Expected integer");
          else if (k == "tunnel") tunnel = v;
        });
      } else if (buf.find("@vars") != std::string::npos) {
        block([&](const std::string& k, const std::string& v) {
          vars.push_back(k);
          vars_map.push_back(v);
        });
      } else if (buf.find("@all") != std::string::npos) {
        fill_extras = false;   // the reference assigns load_all's result: only the LAST @all block decides
        block([&](const std::string& k, const std::string& v) {
          if (k == "sds") sds = v;
          else if (k == "grid") {
            grid = v;
            const std::string frmt = grid.substr(grid.find_last_of('.') + 1);
            if (frmt == "p3d" || frmt == "g" || frmt == "x" || frmt == "grid" || frmt == "grd") grid_type = GridType::P3D;
            if (frmt == "tri") grid_type = GridType::Tri;
          } else if (k == "normals") normals = v;
          else if (k == "targets") { fill_targets = v; fill_extras = true; }
          else if (k == "calibration") { fill_calibration = v; fill_extras = true; }
          else if (k == "grid_units") grid_units = v;
          else if (k == "active_comps") active_comps = v;
        });
      } else if (buf.find("@camera") != std::string::npos) {
        ++cameras;
        cam_nums.push_back(0);
        camera_filenames.push_back("");
        cals.push_back("");
        targets.push_back("");
        block([&](const std::string& k, const std::string& v) {
          if (k == "number") {
            int n = 0;
            if (!to_int(v, n, "")) throw std::invalid_argument("Could not parse camera number in input file");
            cam_nums.back() = (unsigned)n;
          } else if (k == "cine" || k == "filename") camera_filenames.back() = v;
          else if (k == "calibration") cals.back() = v;
          else if (k == "targets") targets.back() = v;
        });
      } else if (buf.find("@options") != std::string::npos) {
        block([&](const std::string& k, const std::string& v) {
          if (k == "target_patcher") {
            if (v == "polynomial") target_patcher = TargetPatchType::Polynomial;
            else if (v == "none") target_patcher = TargetPatchType::None;
            else ok = fail("Error: Could not parse @options:target_patcher. Options are 'polynomial' or 'none'");
          } else if (k == "registration") {
            if (v == "pixel") registration = RegistrationType::Pixel;
            else if (v == "point") registration = RegistrationType::Point;
            else if (v == "none") registration = RegistrationType::None;
            else ok = fail("Error: Could not parse @options:registration. Options are 'pixel', 'point', or 'none'");
          } else if (k == "pixel_interpolation") {
            if (v == "linear") pixel_interpolation = PixelInterpolationType::Linear;
            else if (v == "nearest") pixel_interpolation = PixelInterpolationType::Nearest;
            else ok = fail("Error: Could not parse @options:pixel_interpolation. Options are 'linear' or 'nearest.");
          } else if (k == "filter") {
            if (v == "gaussian") filter = FilterType::Gaussian;
            else if (v == "box") filter = FilterType::Box;
            else if (v == "none") filter = FilterType::None;
            else ok = fail("Error: Could not parse @options:filter. Options are 'gaussian', 'box', or 'none'");
          } else if (k == "overlap") {
            if (v == "best_view") overlap = OverlapKind::BestView;
            else if (v == "average_view") overlap = OverlapKind::AverageViews;
            else ok = fail("Error: Could not parse @options:overlap. Options are 'best_view' or 'average_view'");
          } else if (k == "filter_size") ok = ok && to_int(v, filter_size, "Error: Could not parse @options:filter_size. Expected integer");
          else if (k == "oblique_angle") {
            try { oblique_angle = std::stof(v); } catch (...) { ok = fail("Error: Could not parse @options:oblique_angle. Expected float"); }
          } else if (k == "number_frames") ok = ok && to_int(v, number_frames, "Error: Could not parse @options:number_frames. Expected integer");
        });
      } else if (buf.find("@output") != std::string::npos) {
        block([&](const std::string& k, const std::string& v) {
          if (k == "dir") out_dir = v;
          else if (k == "name") out_name = v;
        });
      }
    }
    if (!ok) return false;
    if (!std::is_sorted(cam_nums.begin(), cam_nums.end())) {      // cameras by number, ascending
      std::vector<size_t> idx(cameras);
      std::iota(idx.begin(), idx.end(), 0);
      std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return cam_nums[a] < cam_nums[b]; });
      auto perm = [&](auto& v) {
        auto t = v;
        for (size_t k = 0; k < idx.size(); ++k) v[k] = t[idx[k]];
      };
      perm(cam_nums); perm(camera_filenames); perm(targets); perm(cals);
    }
    if (!vars.empty()) {
      if (!evaluate_vars(fill_targets)) return fail("Unable to parse variables in @all:targets");
      if (!evaluate_vars(fill_calibration)) return fail("Unable to parse variables in @all:calibration");
      if (!evaluate_vars(sds)) return fail("Unable to parse variables in @all:sds");
      if (!evaluate_vars(grid)) return fail("Unable to parse variables in @all:grid");
      if (!evaluate_vars(normals)) return fail("Unable to parse variables in @all:normals");
      if (!evaluate_vars(active_comps)) return fail("Unable to parse variables in @all:active_comps");
      if (!evaluate_vars(out_dir)) return fail("Unable to parse variables in @output:dir");
      for (unsigned c = 0; c < cameras; ++c) {
        if (!evaluate_vars(targets[c])) return fail("Unable to parse variables in @camera:" + std::to_string(c + 1) + ":targets");
        if (!evaluate_vars(cals[c])) return fail("Unable to parse variables in @camera:" + std::to_string(c + 1) + ":calibration");
        if (!evaluate_vars(camera_filenames[c])) return fail("Unable to parse variables in @camera:" + std::to_string(c + 1) + ":filename");
      }
    }
    if (fill_extras)
      for (unsigned c = 0; c < cameras; ++c) {
        if (targets[c].empty()) targets[c] = fill_targets;
        if (cals[c].empty()) cals[c] = fill_calibration;
      }
    return true;
  }

  bool has_normals() const { return !normals.empty(); }

  /* every file the deck names must exist, the grid type must be known (upsp_inputs.cpp:176-228) */
  bool check_all() {
    struct stat st;
    auto missing = [&](const std::string& what, const std::string& path) {
      if (stat(path.c_str(), &st) == 0) return false;
      fail(what + " '" + path + "': " + std::strerror(errno));
      return true;
    };
    for (unsigned c = 0; c < cameras; ++c) {
      if (missing("Camera " + std::to_string(c + 1) + " file", camera_filenames[c])) return false;
      if (missing("Targets file", targets[c])) return false;
      if (missing("Calibration file", cals[c])) return false;
    }
    if (missing("SDS file", sds)) return false;
    if (missing("Grid file", grid)) return false;
    if (grid_type == GridType::None) return fail("Invalid GridType::None");
    if (!normals.empty() && missing("Normals file", normals)) return false;
    if (!active_comps.empty() && missing("Active components file", active_comps)) return false;
    if (missing("Output dir", out_dir)) return false;
    return true;
  }

  /* the deck written back (upsp_inputs.cpp:236-341): values that contain a variable's value get `$name` again
   * (the longest matching value wins), targets / calibration move to @all when every camera shares them.
   * `date` fills %Date_Created (the reference prints today's m/d/yyyy). */
  void write_file(const std::string& out_file, const std::string& date = "") const;
  std::string refill_vars(std::string term) const {
    if (vars.empty()) return term;
    int best_match = -1;
    size_t best_size = 0;
    for (size_t i = 0; i < vars.size(); ++i)
      if (term.find(vars_map[i]) != std::string::npos && vars_map[i].size() > best_size) {
        best_match = (int)i;
        best_size = vars_map[i].size();
      }
    if (best_match != -1)
      term.replace(term.find(vars_map[(size_t)best_match]), vars_map[(size_t)best_match].length(), "$" + vars[(size_t)best_match]);
    return term;
  }

  /* $name substitution, one pass per '$' in the term (upsp_inputs.cpp:668-699) */
  bool evaluate_vars(std::string& term) const {
    if (term.empty()) return true;
    const size_t num_vars = (size_t)std::count(term.begin(), term.end(), '$');
    for (size_t j = 0; j < num_vars; ++j) {
      bool replaced = false;
      for (size_t i = 0; i < vars.size(); ++i) {
        const std::string var = "$" + vars[i];
        const size_t pos_start = term.find(var);
        if (pos_start != std::string::npos) {
          term.replace(pos_start, pos_start + var.size(), vars_map[i]);    // the reference's call, as is
          replaced = true;
          break;
        }
      }
      if (!replaced) return false;
    }
    return term.find('$') == std::string::npos;
  }

 private:
  bool fail(const std::string& msg) {
    error = msg;
    return false;
  }
  bool to_int(const std::string& v, int& out, const std::string& msg) {
    try {
      out = std::stoi(v);
      return true;
    } catch (...) {
      return msg.empty() ? false : fail(msg);
    }
  }
  static std::vector<std::string> parse_line(std::string input) {
    input.erase(std::remove_if(input.begin(), input.end(), [](unsigned char c) { return std::isspace(c); }), input.end());
    std::vector<std::string> tokens;
    size_t a = 0;
    for (size_t p; (p = input.find('=', a)) != std::string::npos; a = p + 1) tokens.push_back(input.substr(a, p - a));
    if (a < input.size()) tokens.push_back(input.substr(a));   // std::getline(ss, seg, '='): no piece after a trailing '='
    return tokens;
  }
  static std::vector<std::string> split_whitespace(const std::string& s) {
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) {
      if (std::isspace((unsigned char)c)) {
        if (!cur.empty()) out.push_back(cur), cur.clear();
      } else cur += c;
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
  }
};

inline const char* to_string(TargetPatchType v) { return v == TargetPatchType::Polynomial ? "polynomial" : "none"; }
inline const char* to_string(RegistrationType v) {
  return v == RegistrationType::Pixel ? "pixel" : (v == RegistrationType::Point ? "point" : "none");
}
inline const char* to_string(PixelInterpolationType v) { return v == PixelInterpolationType::Nearest ? "nearest" : "linear"; }
inline const char* to_string(FilterType v) { return v == FilterType::Gaussian ? "gaussian" : (v == FilterType::Box ? "box" : "none"); }
inline const char* to_string(OverlapKind v) { return v == OverlapKind::BestView ? "best_view" : "average_view"; }
/* to_string: the deck keywords (what Load accepts).  display_name: what the reference's stream operators print
 * (upsp_inputs.cpp:802-882), which differ in two places: a box filter prints as "undefined filter type" (no case for it)
 * and the averaging overlap as "average_views" (the keyword is "average_view"); its write_file inherits both. */
inline const char* display_name(TargetPatchType v) { return to_string(v); }
inline const char* display_name(RegistrationType v) { return to_string(v); }
inline const char* display_name(PixelInterpolationType v) { return to_string(v); }
inline const char* display_name(FilterType v) { return v == FilterType::Box ? "undefined filter type" : to_string(v); }
inline const char* display_name(OverlapKind v) { return v == OverlapKind::BestView ? "best_view" : "average_views"; }
inline const char* to_string(GridType v) { return v == GridType::P3D ? "p3d" : (v == GridType::Tri ? "tri" : "none"); }

inline void FileInputs::write_file(const std::string& out_file, const std::string& date) const {
  std::ofstream ofs(out_file);
  if (!ofs.is_open()) throw std::invalid_argument("Could not open file for writing input deck");
  if (!version.empty()) ofs << "%Version " << version << "\n%Date_Created " << date << "\n\n";
  ofs << "@general\n\ttest = " << test_id << "\n\trun = " << run << "\n\tsequence = " << sequence << "\n";
  if (!tunnel.empty()) ofs << "\ttunnel = " << tunnel << "\n";
  if (!grid_units.empty()) ofs << "\tgrid_units = " << grid_units << "\n";
  if (!vars.empty()) {
    ofs << "@vars\n";
    for (size_t i = 0; i < vars.size(); ++i) ofs << "\t" << vars[i] << " = " << vars_map[i] << "\n";
  }
  ofs << "@all\n\tgrid = " << refill_vars(grid) << "\n\tsds = " << refill_vars(sds) << "\n";
  if (!normals.empty()) ofs << "\tnormals = " << refill_vars(normals) << "\n";
  if (!active_comps.empty()) ofs << "\tactive_comps = " << refill_vars(active_comps) << "\n";
  const bool targ_all = std::all_of(targets.begin(), targets.end(), [&](const std::string& t) { return t == targets[0]; });
  const bool cal_all = std::all_of(cals.begin(), cals.end(), [&](const std::string& t) { return t == cals[0]; });
  if (targ_all && !targets.empty()) ofs << "\ttargets = " << refill_vars(targets[0]) << "\n";
  if (cal_all && !cals.empty()) ofs << "\tcalibration = " << refill_vars(cals[0]) << "\n";
  for (unsigned c = 0; c < cameras; ++c) {
    ofs << "@camera\n\tnumber = " << cam_nums[c] << "\n\tfilename = " << refill_vars(camera_filenames[c]) << "\n";
    if (!targ_all) ofs << "\ttargets = " << refill_vars(targets[c]) << "\n";
    if (!cal_all) ofs << "\tcalibration = " << refill_vars(cals[c]) << "\n";
  }
  ofs << "@options\n\ttarget_patcher = " << display_name(target_patcher) << "\n\tregistration = " << display_name(registration)
      << "\n\tpixel_interpolation = " << display_name(pixel_interpolation) << "\n\tfilter = " << display_name(filter)
      << "\n\toverlap = " << display_name(overlap) << "\n\tfilter_size = " << filter_size << "\n\toblique_angle = " << oblique_angle
      << "\n\tnumber_frames = " << number_frames << "\n";
  ofs << "@output\n\tdir = " << refill_vars(out_dir) << "\n\tname = " << out_name << "\n";
}

/* the summary psp_process prints on rank 0 (upsp_inputs.cpp:753-798) */
inline std::ostream& operator<<(std::ostream& os, const FileInputs& fi) {
  os << "\nInput File " << fi.filename << ":\n" << fi.test_id << " Run " << fi.run << " Seq " << fi.sequence << "\n" << fi.tunnel << "\n\n";
  os << "grid         = " << fi.grid << "\n    units    = " << fi.grid_units << "\n";
  if (!fi.normals.empty()) os << "normals      = " << fi.normals << "\n";
  if (!fi.active_comps.empty()) os << "active_comps = " << fi.active_comps << "\n";
  os << "sds          = " << fi.sds << "\n\n" << fi.cameras << (fi.cameras > 1 ? " Cameras:\n" : " Camera:\n");
  for (unsigned c = 0; c < fi.cameras; ++c)
    os << " camera " << fi.cam_nums[c] << "\n           filename    = " << fi.camera_filenames[c] << "\n           targets     = "
       << fi.targets[c] << "\n           calibration = " << fi.cals[c] << "\n";
  os << "\noutput dir  = " << fi.out_dir << "\noutput name = " << fi.out_name << "\n\nOptions:\n";
  os << "  Target Patcher   = " << display_name(fi.target_patcher) << "\n  Registration     = " << display_name(fi.registration)
     << "\n  Pixel Interp     = " << display_name(fi.pixel_interpolation) << "\n  Filter           = " << display_name(fi.filter) << " ("
     << fi.filter_size << "x" << fi.filter_size << ")\n  Overlap          = " << display_name(fi.overlap)
     << "\n  Oblique Angle    = " << fi.oblique_angle << "\n  Number of Frames = " << fi.number_frames << "\n";
  return os;
}

}  // namespace upsp_b200
