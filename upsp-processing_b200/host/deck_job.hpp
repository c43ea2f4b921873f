// deck_job.hpp -- psp_process's start-up from its own command line and input deck to a job directory of the frame
// chain (shared by psp_setup_b200 and by psp_process_b200's one-step mode).  See psp_setup_b200.cpp for the option list
// and what each step mirrors (ParseOpts cpp/exec/psp_process.cpp:1192-1310, InitializeVideoStreams :392-470,
// InitializeModel :2183-2196, InitializeProjection :1595-1641, phase-2 start-up :2270-2385).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <map>

#include "camera_cal.hpp"
#include "grid_readers.hpp"
#include "interpolation.hpp"
#include "p3d_model.hpp"
#include "projection_weights.hpp"
#include "run_inputs.hpp"
#include "targets.hpp"
#include "upsp_inputs.hpp"
#include "video_readers.hpp"

namespace upsp_b200 {

/* the reference's cv::CommandLineParser takes -key=value; -key value is accepted as well.  Flags without a value: see `bare` */
inline std::map<std::string, std::string> parse_command_line(int argc, char** argv, const std::vector<std::string>& bare = {"-no_projection"}) {
  std::map<std::string, std::string> opt;
  for (int k = 1; k < argc; ++k) {
    std::string a = argv[k];
    const size_t eq = a.find('=');
    if (a.size() > 1 && a[0] == '-' && eq != std::string::npos) {
      opt[a.substr(0, eq)] = a.substr(eq + 1);
      continue;
    }
    if (std::find(bare.begin(), bare.end(), a) != bare.end()) opt[a] = "1";
    else if (k + 1 < argc) opt[a] = argv[++k];
    else throw std::invalid_argument("missing value after " + a);
  }
  return opt;
}

template <typename T>
inline void write_all(const std::string& path, const std::vector<T>& v) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::invalid_argument("Cannot write '" + path + "'");
  std::fwrite(v.data(), sizeof(T), v.size(), f);
  std::fclose(f);
}

struct ModelArrays {
  std::vector<float> xyz, normals;      // [N][3]; normals = Model::get_n() (create_projection_mat, getTargets)
  std::vector<float> node_normals;      // [N][3]  Node::get_normal() (camera weights, target diameters): the same array for a
                                        // structured model, the area-weighted one for an unstructured model
  std::vector<int32_t> tris;            // [T][3]
  std::vector<uint8_t> is_data;
  std::vector<int32_t> remap;           // structured grids only
  std::vector<int32_t> primary_comp;    // per node: its component, or INT32_MIN when it has none (Node::has_primary_component)
  std::vector<int32_t> tri_comps;       // [T] component of every triangle (unstructured; 0 where the file has none): /Grid/components
  std::vector<int32_t> zone_sizes;      // [zones][3] (structured): /Grid/grid_sizes
  int n_components = 0;
  bool structured = false;
  int n_nodes() const { return (int)(xyz.size() / 3); }
};

inline ModelArrays load_model(const FileInputs& ifile) {
  ModelArrays m;
  if (ifile.grid_type == GridType::Tri) {
    TriGrid g = read_tri_grid(ifile.grid);
    std::cout << "Read " << g.n_nodes << " nodes and " << g.n_tris << " faces" << std::endl;
    {   // TriModel_(file, intersect = true), psp_process.cpp:1384 / TriModel.ipp:245-255
      std::cout << "Finding overlapping points..." << std::endl;
      const int initial_size = g.n_nodes;
      const int overlap = intersect_grid(g);
      std::cout << "Found " << (initial_size - g.n_nodes) << " non-unique points\nFound " << overlap << " unique overlapping points\n" << std::endl;
    }
    m.xyz = g.xyz;
    m.tris = g.tris;
    m.tri_comps = g.comps;
    m.tri_comps.resize((size_t)g.n_tris, 0);
    calc_normals(g, m.normals);
    node_normals_area_weighted(g, m.node_normals);
    m.is_data.assign((size_t)g.n_nodes, 1);
    // TriModel_::Node::has_primary_component (TriModel.ipp:1530-1546): every adjacent triangle carries the same component
    m.n_components = g.number_of_components();
    m.primary_comp.assign((size_t)g.n_nodes, INT32_MIN);
    if (!g.comps.empty()) {
      std::vector<uint8_t> seen((size_t)g.n_nodes, 0);
      for (int t = 0; t < g.n_tris; ++t)
        for (int k = 0; k < 3; ++k) {
          const size_t n = (size_t)g.tris[(size_t)t * 3 + k];
          if (!seen[n]) seen[n] = 1, m.primary_comp[n] = g.comps[(size_t)t];
          else if (seen[n] == 1 && m.primary_comp[n] != g.comps[(size_t)t]) seen[n] = 2, m.primary_comp[n] = INT32_MIN;
        }
    }
  } else {
    const P3DModel model(ifile.grid, 1e-3f);            // psp_process.cpp:1378
    m.structured = true;
    const int N = model.size();
    m.xyz.resize((size_t)N * 3);
    for (int n = 0; n < N; ++n) {
      m.xyz[(size_t)n * 3 + 0] = model.get_x()[(size_t)n];
      m.xyz[(size_t)n * 3 + 1] = model.get_y()[(size_t)n];
      m.xyz[(size_t)n * 3 + 2] = model.get_z()[(size_t)n];
    }
    m.normals = model.get_n();
    m.node_normals = m.normals;           // P3DModel_::Node::get_normal returns the stored normal (P3DModel.ipp:1685-1687)
    std::vector<float> unused;
    std::vector<int> tn;
    model.extract_tris(unused, tn);
    m.tris.assign(tn.begin(), tn.end());
    m.is_data.resize((size_t)N);
    for (int n = 0; n < N; ++n) m.is_data[(size_t)n] = model.is_superceded(n) ? 0 : 1;   // the node iterator skips them
    m.remap = model.overlap_src_index();
    m.n_components = model.num_zones();                   // P3DModel.h:244, Node::get_primary_component = zone
    for (int zn = 0; zn < model.num_zones(); ++zn) {        // PSPHDF5.ipp:170-176: zone_size(i, 0), (i, 1), (i, 2)
      m.zone_sizes.push_back(model.zone_size(zn, 0));
      m.zone_sizes.push_back(model.zone_size(zn, 1));
      m.zone_sizes.push_back(1);
    }
    m.primary_comp.resize((size_t)N);
    for (int n = 0; n < N; ++n) m.primary_comp[(size_t)n] = model.nidx2_gidx(n).zone;
  }
  return m;
}

inline int run_deck(const std::map<std::string, std::string>& opt) {
  auto has = [&](const char* k) { return opt.count(k) != 0; };
  auto get = [&](const char* k) { return opt.at(k); };
  auto fail = [](const std::string& msg) {
    std::cerr << "[ERROR] " << msg << std::endl;
    return 1;
  };
  if (!has("-paint_cal")) return fail("Must specify -paint_cal");
  if (!has("-job_dir")) return fail("Must specify -job_dir");
  FileInputs ifile;
  if (!ifile.Load(get("-input_file"))) return fail(ifile.error);
  if (has("-frames")) ifile.number_frames = std::atoi(get("-frames").c_str());
  if (!ifile.check_all()) return fail(ifile.error);
  if (ifile.tunnel != "ames_unitary") return fail("Unrecognized tunnel name '" + ifile.tunnel + "'");
  if (ifile.registration != RegistrationType::None && ifile.registration != RegistrationType::Pixel)
    return fail("Unsupported registration type");
  if (ifile.filter_size % 2 == 0) return fail("Filter size must be odd (currently '" + std::to_string(ifile.filter_size) + "')");
  std::cout << ifile << std::endl;
  const std::string job_dir = get("-job_dir");
  const int device = has("-device") ? std::atoi(get("-device").c_str()) : 0;

  // ---- video streams, number of frames (InitializeVideoStreams) ----
  unsigned number_frames = ifile.number_frames < 0 ? std::numeric_limits<unsigned>::max() : (unsigned)ifile.number_frames;
  std::vector<std::unique_ptr<VideoReader>> cams(ifile.cameras);
  for (unsigned c = 0; c < ifile.cameras; ++c) {
    const std::string& fn = ifile.camera_filenames[c];
    const size_t dot = fn.rfind('.');
    const std::string ext = dot == std::string::npos ? "" : fn.substr(dot);
    if (ext != ".cine" && ext != ".mraw")
      return fail("Unknown video file extension '" + ext + "' for '" + fn + "'. Valid extensions: {'.cine', '.mraw'}");
    cams[c] = open_video(fn);
    const VideoProperties& vp = cams[c]->properties();
    if (ifile.number_frames < 0) number_frames = std::min(number_frames, vp.num_frames);
    else if (vp.num_frames < number_frames)
      return fail("(" + std::to_string(number_frames) + ") frames requested but only (" + std::to_string(vp.num_frames) +
                  ") frames available in '" + fn + "'");
    std::cout << "Initialized video stream ['" << fn << "']\n  Frames per second : " << vp.frame_rate << "\n  Frame size        : ["
              << vp.width << " x " << vp.height << "]\n  Bit depth         : " << vp.bit_depth << std::endl;
    if (vp.width != cams[0]->properties().width || vp.height != cams[0]->properties().height)
      return fail("cameras with different frame sizes are not supported by psp_process_b200");
  }
  std::cout << "Will process (" << number_frames << ") frames" << std::endl;
  const int W = (int)cams[0]->properties().width, H = (int)cams[0]->properties().height;

  // ---- model ----
  ModelArrays model = load_model(ifile);
  const int msize = model.n_nodes();
  std::cout << "Loaded model: " << msize << " nodes, " << model.tris.size() / 3 << " triangles" << std::endl;
  if (ifile.has_normals()) {                             // InitializeModel, psp_process.cpp:2185-2189
    if (!model.structured)
      std::cerr << "[ERROR] Refusing to read '" << ifile.normals << "'; can not specify normals CSV for a TriModel_" << std::endl;
    else
      std::cout << "Overwrote " << set_surface_normals(ifile.normals, model.normals) << "/" << msize << " model surface normals (using '"
                << ifile.normals << "')" << std::endl;
  }
  if (!ifile.active_comps.empty()) {                     // psp_process.cpp:1462-1486
    const auto active = read_active_comp_file(ifile.active_comps);
    if ((int)active.size() > model.n_components)
      return fail("Error: Number of components in active component file cannot be greater than the number of components in the grid");
    for (int n = 0; n < msize; ++n) {
      if (!model.is_data[(size_t)n] && model.structured) continue;      // the node iterator skips superceded nodes
      const int32_t comp = model.primary_comp[(size_t)n];
      if (comp == INT32_MIN) continue;
      const auto it = active.find(comp);
      if (it != active.end() && !it->second) model.is_data[(size_t)n] = 0;
    }
  }
  if (has("-cutoff_x_max")) {
    const float x_max = (float)std::atof(get("-cutoff_x_max").c_str());
    for (int n = 0; n < msize; ++n)
      if (model.xyz[(size_t)n * 3] > x_max) model.is_data[(size_t)n] = 0;
  }

  // ---- phase-2 inputs ----
  const PaintCalibration pcal(get("-paint_cal"));
  TunnelConditions tcond = read_tunnel_conditions(ifile.sds);
  float wall_temp = 0.f;
  const float model_temp = model_temperature(tcond, &wall_temp);
  if (!std::isnan(tcond.tcavg))
    std::cout << "*** Using thermocouple average (" << tcond.tcavg << "F) for model temp, supersedes estimated temperature based on "
              << "boundary layer recovery factor (" << wall_temp << "F)" << std::endl;
  else
    std::cout << "*** Using estimated temperature based on boundary layer recovery factor (" << wall_temp << "F)" << std::endl;
  std::vector<float> model_temp_input((size_t)msize, model_temp), steady((size_t)msize, 0.0f);
  auto read_function = [&](const char* key, const char* what, std::vector<float>& dst) -> bool {
    if (!has(key) || get(key).empty()) return true;
    if (!model.structured) {   // psp_process.cpp:2338-2345, 2371-2378: k-nearest inverse-distance interpolation from the steady grid
      if (!has("-steady_grid") || get("-steady_grid").empty()) {
        fail(std::string(what) + " function file with an unstructured grid needs -steady_grid");
        return false;
      }
      const std::vector<float> in = read_plot3d_scalar_function_file(get(key));
      const P3DModel steady_grid(get("-steady_grid"), 1e-3f);
      if ((int)in.size() != steady_grid.size()) {
        fail(std::string(what) + " function file inconsistent with the steady grid (expect " + std::to_string(steady_grid.size()) +
             " values, got " + std::to_string(in.size()) + ")");
        return false;
      }
      dst = interpolate(steady_grid, in, model.xyz.data(), msize, 10, 2.0f);
      return true;
    }
    dst = read_plot3d_scalar_function_file(get(key));
    if ((int)dst.size() != msize) {
      fail(std::string(what) + " function file inconsistent with grid (expect " + std::to_string(msize) + " values, got " +
           std::to_string(dst.size()) + ")");
      return false;
    }
    return true;
  };
  if (!read_function("-model_temp_p3d", "Model-temperature", model_temp_input)) return 1;
  if (!read_function("-steady_p3d", "Steady-state", steady)) return 1;

  // ---- per camera: calibration + projection matrix ----
  const bool project = !has("-no_projection");
  const float obliqueThresh = (float)((180. - ifile.oblique_angle) * 3.14159265358979323846 / 180.0);
  std::vector<CsrMatrix> projs(ifile.cameras);
  std::vector<std::array<double, 3>> centers(ifile.cameras);
  for (unsigned c = 0; c < ifile.cameras; ++c) {
    const upsp_camera_model cam = read_json_camera_calibration(ifile.cals[c]);
    centers[c] = get_cam_center(cam);
    if (cam.width != W || cam.height != H)
      std::cout << "Warning: calibration imageSize " << cam.width << "x" << cam.height << " differs from the video frames " << W << "x" << H
                << std::endl;
    if (ifile.target_patcher == TargetPatchType::Polynomial) {
      // InitializeImagePatches up to the clustering (psp_process.cpp:2095-2123); psp_process_b200 clusters the projected
      // targets, thresholds the boundaries on the first frame and builds the pixel lists (host/patch_geometry.hpp)
      const float sf = has("-target_diam_sf") ? (float)std::atof(get("-target_diam_sf").c_str()) : 1.2f;
      const std::vector<Target> targs = visible_targets(HostCamera(cam), model.xyz.data(), model.normals.data(), model.node_normals.data(), msize, model.tris.data(),
                                                        (int)(model.tris.size() / 3), ifile.targets[c], ifile.oblique_angle, sf);
      FILE* tf = std::fopen((job_dir + "/cam" + std::to_string(c) + ".targets").c_str(), "w");
      if (!tf) return fail("Cannot write the projected targets of camera " + std::to_string(c + 1));
      for (const Target& t : targs) std::fprintf(tf, "%.9g %.9g %.9g\n", (double)t.u, (double)t.v, (double)t.diameter);
      std::fclose(tf);
      std::cout << "camera " << ifile.cam_nums[c] << ": " << targs.size() << " visible targets / fiducials" << std::endl;
    }
    if (!project) continue;
    std::vector<int32_t> code((size_t)msize);
    std::vector<float> uv((size_t)2 * msize);
    if (upsp_op_create_projection(device, &cam, model.xyz.data(), model.normals.data(), model.is_data.data(), msize, model.tris.data(),
                                  (int)(model.tris.size() / 3), obliqueThresh, code.data(), uv.data()) != UPSP_OK)
      throw std::runtime_error(upsp_gpu_last_error());
    CsrMatrix& m = projs[c];
    m.rowptr.assign(1, 0);
    for (int n = 0; n < msize; ++n) {
      if (code[(size_t)n] >= 0) m.col.push_back(code[(size_t)n]);
      m.rowptr.push_back((int32_t)m.col.size());
    }
    m.val.assign(m.col.size(), 1.0f);
    char name[64];
    std::snprintf(name, sizeof name, "/cam%02u-uv", c + 1);
    write_all((has("-uv_dir") ? get("-uv_dir") : job_dir) + name, uv);      // psp_process.cpp:1615-1620: add_out_dir/camNN-uv
    std::cout << "camera " << ifile.cam_nums[c] << ": projected " << msize << " model nodes, accepted " << m.col.size() << std::endl;
  }
  if (project) {
    adjust_projection_for_weights(model.xyz.data(), model.node_normals.data(), centers, projs,
                                  ifile.overlap == OverlapKind::BestView ? OverlapType::BestView : OverlapType::AverageViews);
    for (unsigned c = 0; c < ifile.cameras; ++c) {
      const std::string b = job_dir + "/cam" + std::to_string(c);
      write_all(b + ".rowptr", projs[c].rowptr);
      write_all(b + ".col", projs[c].col);
      write_all(b + ".val", projs[c].val);
    }
  }

  // ---- job directory ----
  if (model.structured) write_all(job_dir + "/remap.i32", model.remap);
  write_all(job_dir + "/steady.f32", steady);
  write_all(job_dir + "/model_temp.f32", model_temp_input);
  write_all(job_dir + "/xyz.f32", model.xyz);
  write_all(job_dir + "/normals.f32", model.normals);
  write_all(job_dir + "/node_normals.f32", model.node_normals);
  write_all(job_dir + "/is_data.u8", model.is_data);
  // what the HDF5 outputs need beyond the flat files (host/psp_hdf5.hpp, psp_process.cpp:2400-2420)
  write_all(job_dir + "/tris.i32", model.tris);
  if (model.structured) write_all(job_dir + "/grid_sizes.i32", model.zone_sizes);
  else write_all(job_dir + "/tri_comps.i32", model.tri_comps);
  std::ofstream job(job_dir + "/job.txt");
  if (!job) return fail("Cannot write '" + job_dir + "/job.txt'");
  job << std::setprecision(9);
  job << "# written by psp_setup_b200 from " << ifile.filename << "\n";
  job << "cameras = " << ifile.cameras << "\nwidth = " << W << "\nheight = " << H << "\nnumber_frames = " << number_frames
      << "\nmsize = " << msize << "\nfirst_frame = 1\n";
  for (unsigned c = 0; c < ifile.cameras; ++c) job << "video" << c << " = " << ifile.camera_filenames[c] << "\n";
  job << "registration = " << to_string(ifile.registration) << "\npixel_interpolation = " << to_string(ifile.pixel_interpolation)
      << "\ntarget_patcher = " << to_string(ifile.target_patcher) << "\nfilter = " << to_string(ifile.filter)
      << "\nfilter_size = " << ifile.filter_size << "\n";
  job << "bound_thickness = " << (has("-bound_pts") ? get("-bound_pts") : "2") << "\nbuffer_thickness = "
      << (has("-buffer_pts") ? get("-buffer_pts") : "1") << "\nauto_patch_thresh = 1\n";
  job << "qbar = " << tcond.qbar << "\nps = " << tcond.ps << "\ndegree = 6\n";
  job << "cal_a = " << pcal.a << "\ncal_b = " << pcal.b << "\ncal_c = " << pcal.c << "\ncal_d = " << pcal.d << "\ncal_e = " << pcal.e
      << "\ncal_f = " << pcal.f << "\n";
  job << "test_id = " << ifile.test_id << "\nrun = " << ifile.run << "\nsequence = " << ifile.sequence << "\ngrid_units = "
      << ifile.grid_units << "\nout_dir = " << ifile.out_dir << "\nout_name = " << ifile.out_name << "\n";
  {
    // /Condition of the HDF5 files: tunnel conditions (.wtd) + camera settings of camera 1 (psp_process.cpp:1581-1589;
    // get_focal_length = cameraMatrix(0,0) * pix_sz_, and the JSON calibration reader leaves pix_sz_ at -1: CameraCal.cpp:56, 240-249)
    const auto& vp0 = cams[0]->properties();
    job << "structured = " << (model.structured ? 1 : 0) << "\n";
    char line[512];
    std::snprintf(line, sizeof line, "tc_alpha = %.9g\ntc_beta = %.9g\ntc_phi = %.9g\ntc_mach = %.9g\ntc_rey = %.9g\ntc_ptot = %.9g\n"
                  "tc_ttot = %.9g\ntc_tcavg = %.9g\n", (double)tcond.alpha, (double)tcond.beta, (double)tcond.phi, (double)tcond.mach,
                  (double)tcond.rey, (double)tcond.ptot, (double)tcond.ttot, (double)tcond.tcavg);
    job << line;
    std::snprintf(line, sizeof line, "cam_frame_rate = %d\ncam_fstop = %.9g\ncam_exposure = %.9g\n", (int)vp0.frame_rate,
                  (double)vp0.aperture, (double)vp0.exposure);
    job << line;
    job << "cam_focal_lengths =";
    for (unsigned c = 0; c < ifile.cameras; ++c) {
      const upsp_camera_model cam = read_json_camera_calibration(ifile.cals[c]);
      std::snprintf(line, sizeof line, " %.9g", (double)(float)(cam.fx * (double)-1.0f));
      job << line;
    }
    job << "\ncode_version = " << (has("-code_version") ? get("-code_version") : std::string("upsp-processing_b200")) << "\n";
  }
  std::cout << "Wrote job directory " << job_dir << std::endl;
  return 0;
}

}  // namespace upsp_b200
