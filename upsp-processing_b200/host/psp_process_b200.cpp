// psp_process_b200 -- C++ host driver of the GPU frame chain (single rank).
//
// Mirrors the control flow of the reference's psp_process (cpp/exec/psp_process.cpp): phase 1
// (frame loop :1743-1851, reduction / finals :1866-1940, global_transpose :2032) and phase 2
// (node loop :2452-2507, finals :2540-2547), and writes the reference's flat output files
// (:524-540) with the same names and layout.  What it does NOT do is the reference's phase 0
// (input deck, grid readers, camera calibration, BVH visibility -> projection matrix): those
// products are read from a job directory of raw little-endian arrays instead (written by
// upsp-processing_b200/synth.py: write_job, or by any tool that has run phase 0):
//
//   job.txt               key = value lines: cameras width height number_frames msize format
//                         registration pixel_interpolation target_patcher qbar ps cal_a..cal_f degree
//   cam<c>.frames         number_frames frames, u16 or 12-bit packed (format = u16 | p12)
//   video<c> = PATH       (job.txt, instead of cam<c>.frames) a Vision Research .cine or Photron .mraw
//                         file, read with host/video_readers.hpp as the reference's CineReader /
//                         MrawReader do; `first_frame` (1-based, default 1) as the deck's start frame
//   cam<c>.rowptr/.col/.val   projection matrix of camera c in CSR (i32, i32, f32)
//   cam<c>.warp           registration = given : number_frames x 6 f32
//   cam<c>.first          registration = pixel : raw first frame, u16
//   cam<c>.patch_boff/.patch_bx/.patch_by/.patch_ioff/.patch_ix/.patch_iy   target_patcher = polynomial
//   cam<c>.targets        (instead of the six patch_* files) text, one projected target "u v diameter" per
//                         line: the pixel lists are then built here as InitializeImagePatches does
//                         (psp_process.cpp:2125-2163) with host/patch_geometry.hpp; job.txt keys
//                         bound_thickness (2), buffer_thickness (0), patch_thresh (optional: boundary
//                         pixels near values of cam<c>.first below it are dropped, offset 2)
//   auto_patch_thresh = 1 (job.txt) the threshold of the reference instead of patch_thresh: first-frame histogram,
//                         edges[first_min_threshold(counts, 5)] + 5 (psp_process.cpp:2150-2155, host/run_inputs.hpp)
//   filter / filter_size  (job.txt) deck @options filter: none | gaussian | box
//   remap.i32             optional: static form of P3DModel::adjust_solution
//   xyz.f32               optional [msize][3]: written out as the X, Y, Z flat files (psp_process.cpp:1549-1556)
//   A job directory is either assembled by hand / synth.py, or written by `psp_setup_b200 -input_file DECK ...`
//   from the reference's own inputs (deck, grid, calibrations, paint calibration, .wtd, steady-state file).
//   steady.f32 model_temp.f32   [msize]
//
//   psp_process_b200 -job_dir DIR -out_dir DIR [-device 0] [-chunk 256]
//
// One-step mode, the reference's own command line (ParseOpts, psp_process.cpp:1192-1310; -key=value or -key value):
//   psp_process_b200 -input_file=DECK -h5_out=FILE -paint_cal=FILE [-steady_p3d=FILE] [-steady_grid=FILE]
//                    [-model_temp_p3d=FILE] [-frames=N] [-add_out_dir=DIR] [-bound_pts=2] [-buffer_pts=1]
//                    [-target_diam_sf=1.2] [-cutoff_x_max=X] [-checkout=T] [-device 0] [-chunk 256]
// runs the start-up of host/deck_job.hpp into <add_out_dir>/job_b200 and then the frame chain; flat files go to
// -add_out_dir (default: the deck's @output dir), as in the reference.  -h5_out is required as there; it and
// <add_out_dir>/extras.h5 are written by host/psp_hdf5.hpp (hand-written HDF5 subset: no HDF5 library in this image).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>

#include <sys/stat.h>

#include "psp_hdf5.hpp"
#include "deck_job.hpp"
#include "patch_geometry.hpp"
#include "run_inputs.hpp"
#include "upsp_b200.hpp"
#include "video_readers.hpp"

using namespace upsp_b200;

template <typename T>
static std::vector<T> read_all(const std::string& path, bool required = true) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) {
    if (required) throw std::invalid_argument("Cannot open '" + path + "'");
    return {};
  }
  const std::streamsize n = f.tellg();
  f.seekg(0);
  std::vector<T> v((size_t)n / sizeof(T));
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
  return v;
}

static std::map<std::string, std::string> read_job(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::invalid_argument("Cannot open '" + path + "'");
  std::map<std::string, std::string> kv;
  std::string line;
  while (std::getline(f, line)) {
    const size_t eq = line.find('=');
    if (eq == std::string::npos || line[0] == '#') continue;
    auto trim = [](std::string s) {
      const size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
      return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    };
    kv[trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
  }
  return kv;
}

int main(int argc, char** argv) {
  std::string job_dir, out_dir;
  int device = 0, chunk = 256;
  std::string h5_out;      // -h5_out of the reference's command line (empty: job-directory mode, flat files only)
  try {
    auto opt = parse_command_line(argc, argv, {"-no_projection", "-help", "-h", "-usage", "-?", "--help"});
    if (opt.count("-help") || opt.count("-h") || opt.count("-usage") || opt.count("-?") || opt.count("--help")) {
      // ParseOpts prints the parser's message and main() ends with status 1 (psp_process.cpp:1225-1228, 1362-1367)
      std::cout << "Process Unsteady Pressure-Sensitive Paint (uPSP) video files into pressure-time history on model surface grid\n"
                   "Usage: psp_process_b200 [params]\n"
                   "\t-input_file       full input deck\n"
                   "\t-frames           override input_file number of frames\n"
                   "\t-code_version     version of the repo (deprecated) [XYZ]\n"
                   "\t-trans_nodes      number of nodes per chunk in transposed solution (unused, kept for compatibility) [250]\n"
                   "\t-add_out_dir      output directory for any additional debugging files (default=input deck-specified output directory)\n"
                   "\t-checkout         perform calibration update checkout [F]\n"
                   "\t-bound_pts        thickness of cluster boundary [2]\n"
                   "\t-buffer_pts       thickness of buffer between targets and cluster boundary [1]\n"
                   "\t-target_diam_sf   scale factor to apply onto target diameter [1.2]\n"
                   "\t-cutoff_x_max     ignore nodes beyond this value in projection\n"
                   "\t-h5_out           output hdf5 file\n"
                   "\t-steady_p3d       steady state p3d function file (for wind-on); units of Cp\n"
                   "\t-steady_grid      steady state p3d grid (needed for wind-on unstructured)\n"
                   "\t-paint_cal        unsteady gain paint calibration file\n"
                   "\t-model_temp_p3d   temperature p3d function file; units of Temperature degrees F\n"
                   "\t-device, -chunk   GPU ordinal [0], frames per push / process call [256]\n"
                   "or, from a job directory: psp_process_b200 -job_dir DIR -out_dir DIR [-h5_out FILE]" << std::endl;
      std::cerr << "[ERROR] Failed to validate command line options; aborting" << std::endl;
      return 1;
    }
    if (opt.count("-device")) device = atoi(opt["-device"].c_str());
    if (opt.count("-chunk")) chunk = atoi(opt["-chunk"].c_str());
    if (opt.count("-input_file")) {       // the reference's command line
      if (!opt.count("-h5_out")) {       // ParseOpts' order: input_file, h5_out, paint_cal, then the deck (psp_process.cpp:1230-1249)
        std::cerr << "[ERROR] Must specify -h5_out" << std::endl;
        return 1;
      }
      if (!opt.count("-paint_cal")) {
        std::cerr << "[ERROR] Must specify -paint_cal" << std::endl;
        return 1;
      }
      FileInputs peek;
      if (!peek.Load(opt["-input_file"])) {
        std::cerr << "[ERROR] " << peek.error << std::endl;
        return 1;
      }
      out_dir = opt.count("-add_out_dir") ? opt["-add_out_dir"] : peek.out_dir;
      job_dir = out_dir + "/job_b200";
      mkdir(job_dir.c_str(), 0755);
      opt["-job_dir"] = job_dir;
      opt["-uv_dir"] = out_dir;
      if (run_deck(opt)) return 1;
      if (opt.count("-checkout") && (opt["-checkout"] == "T" || opt["-checkout"] == "true" || opt["-checkout"] == "1")) {
        std::cout << "Checkout complete (calibration / patch start-up only, psp_process.cpp:1576-1579)" << std::endl;
        return 0;
      }
      h5_out = opt["-h5_out"];
    } else {
      if (opt.count("-job_dir")) job_dir = opt["-job_dir"];
      if (opt.count("-out_dir")) out_dir = opt["-out_dir"];
      if (opt.count("-h5_out")) h5_out = opt["-h5_out"];
    }
  } catch (const std::exception& e) {
    std::cerr << "psp_process_b200: " << e.what() << std::endl;
    return 1;
  }
  if (job_dir.empty() || out_dir.empty()) {
    std::cerr << "usage: psp_process_b200 -job_dir DIR -out_dir DIR [-device 0] [-chunk 256]\n"
                 "       psp_process_b200 -input_file=DECK -h5_out=FILE -paint_cal=FILE [psp_process options]" << std::endl;
    return 1;
  }
  try {
    auto job = read_job(job_dir + "/job.txt");
    auto geti = [&](const char* k) { return atoi(job.at(k).c_str()); };
    auto getf = [&](const char* k) { return (float)atof(job.at(k).c_str()); };
    const int cameras = geti("cameras"), W = geti("width"), H = geti("height");
    const int number_frames = geti("number_frames"), msize = geti("msize");
    const std::string reg = job.at("registration"), patcher = job.at("target_patcher");
    // frame sources: raw dumps (cam<c>.frames) or camera files (video<c> = path)
    std::vector<std::unique_ptr<VideoReader>> readers(cameras);
    bool any_video = false;
    for (int c = 0; c < cameras; ++c) {
      auto it = job.find("video" + std::to_string(c));
      if (it == job.end()) continue;
      readers[c] = open_video(it->second.front() == '/' ? it->second : job_dir + "/" + it->second);
      any_video = true;
      const auto& vp = readers[c]->properties();
      if ((int)vp.width != W || (int)vp.height != H)
        throw std::invalid_argument("video" + std::to_string(c) + " is " + std::to_string(vp.width) + "x" +
                                    std::to_string(vp.height) + ", job.txt says " + std::to_string(W) + "x" + std::to_string(H));
    }
    const int first_frame = job.count("first_frame") ? geti("first_frame") : 1;
    for (int c = 0; c < cameras; ++c)
      if (readers[c] && first_frame - 1 + number_frames > (int)readers[c]->properties().num_frames)
        throw std::invalid_argument("video" + std::to_string(c) + " has fewer frames than first_frame + number_frames");
    const bool p12 = job.count("format") ? job.at("format") == "p12" : false;
    if (!any_video && !job.count("format")) throw std::invalid_argument("job.txt: neither format nor video<c> given");
    const size_t frame_bytes = p12 ? (size_t)W * H * 3 / 2 : (size_t)W * H * 2;

    // frame `first_frame` of camera c, decoded.  fix_hot = true: first_frames_raw of the reference (psp_process.cpp:877-884,
    // hot pixels fixed: the registration reference); fix_hot = false: cams[c]->get_frame(1) as InitializeImagePatches takes it
    // for the histogram threshold and threshold_bounds (:2091, 2145-2155), i.e. without the hot-pixel fix
    auto first_frame_of = [&](int c, bool fix_hot) {
      const std::string b = job_dir + "/cam" + std::to_string(c);
      if (!readers[c]) {
        auto first = read_all<uint16_t>(b + ".first");
        if (first.size() != (size_t)W * H) throw std::invalid_argument("cam.first has the wrong size");
        return first;
      }
      std::vector<uint8_t> packed(readers[c]->frame_bytes());
      readers[c]->read_packed((unsigned)first_frame, packed.data());
      std::vector<uint16_t> first((size_t)W * H);
      if (upsp_op_unpack(device, packed.data(), readers[c]->pixel_format(), first.size(), readers[c]->unpack_lut(), first.data()) != UPSP_OK)
        throw std::runtime_error(upsp_gpu_last_error());
      int n_hot = 0;
      if (fix_hot && upsp_op_fix_hot_pixels(device, first.data(), 1, H, W, &n_hot) != UPSP_OK) throw std::runtime_error(upsp_gpu_last_error());
      return first;
    };

    upsp_gpu_config cfg{};
    cfg.device = device;
    cfg.n_cams = cameras;
    cfg.n_nodes = msize;
    cfg.n_frames_total = number_frames;
    cfg.rank = 0;
    cfg.n_ranks = 1;
    cfg.frame_capacity = 2 * chunk;     // streaming ring: the reader stays ahead of the GPU
    cfg.pressure_aliases_intensity = 1;
    FrameChain chain(cfg);

    std::cout << "Initializing projection / patches (phase 0 products from " << job_dir << ")" << std::endl;
    for (int c = 0; c < cameras; ++c) {
      const std::string b = job_dir + "/cam" + std::to_string(c);
      chain.set_camera(c, W, H);
      auto rowptr = read_all<int32_t>(b + ".rowptr"), col = read_all<int32_t>(b + ".col");
      auto val = read_all<float>(b + ".val");
      if ((int)rowptr.size() != msize + 1) throw std::invalid_argument("rowptr size inconsistent with msize");
      if (rowptr.back() < 0 || col.size() != (size_t)rowptr.back() || val.size() != col.size())
        throw std::invalid_argument("job directory: cam" + std::to_string(c) + ".col/.val do not hold rowptr[msize] entries");
      chain.set_projection(c, rowptr.data(), col.data(), val.data());
      std::ifstream targets_file(b + ".targets");
      if (patcher == "polynomial" && targets_file) {
        std::vector<Target> targs;
        Target t;
        while (targets_file >> t.u >> t.v >> t.diameter) targs.push_back(t);
        const unsigned bt = job.count("bound_thickness") ? (unsigned)geti("bound_thickness") : 2u;
        const unsigned bf = job.count("buffer_thickness") ? (unsigned)geti("buffer_thickness") : 0u;
        std::vector<std::vector<Target>> clusters;
        cluster_points(targs, clusters, (int)(bt + bf));
        PatchClusters pc(clusters, W, H, bt, bf);
        if (job.count("patch_thresh") || (job.count("auto_patch_thresh") && geti("auto_patch_thresh"))) {
          const auto first = first_frame_of(c, false);
          const unsigned bit_depth = readers[c] ? readers[c]->properties().bit_depth : 12u;
          const unsigned thresh = job.count("patch_thresh") ? (unsigned)geti("patch_thresh")
                                                            : patch_threshold(first.data(), first.size(), bit_depth);
          pc.threshold_bounds(first.data(), thresh, 2u);
        }
        std::vector<int32_t> boff, ioff;
        std::vector<uint32_t> bx, by, ix, iy;
        pc.flatten(boff, bx, by, ioff, ix, iy);
        std::cout << "Sorted " << targs.size() << " targets into " << clusters.size() << " clusters" << std::endl;
        chain.set_patches(c, (int)clusters.size(), boff.data(), bx.data(), by.data(), ioff.data(), ix.data(), iy.data());
      } else if (patcher == "polynomial") {
        auto boff = read_all<int32_t>(b + ".patch_boff"), ioff = read_all<int32_t>(b + ".patch_ioff");
        auto bx = read_all<uint32_t>(b + ".patch_bx"), by = read_all<uint32_t>(b + ".patch_by");
        auto ix = read_all<uint32_t>(b + ".patch_ix"), iy = read_all<uint32_t>(b + ".patch_iy");
        if (boff.empty() || boff.size() != ioff.size() || (size_t)boff.back() != bx.size() || bx.size() != by.size() ||
            (size_t)ioff.back() != ix.size() || ix.size() != iy.size())
          throw std::invalid_argument("job directory: patch arrays of camera " + std::to_string(c) + " do not match their offsets");
        chain.set_patches(c, (int)boff.size() - 1, boff.data(), bx.data(), by.data(), ioff.data(),
                          ix.data(), iy.data());
      }
      if (reg == "pixel") chain.set_reference_frame(c, first_frame_of(c, true).data());
    }
    auto remap = read_all<int32_t>(job_dir + "/remap.i32", false);
    if (!remap.empty()) chain.set_overlap_remap(remap.data());
    const int regmode = reg == "pixel" ? UPSP_REG_PIXEL : (reg == "given" ? UPSP_REG_GIVEN : UPSP_REG_NONE);
    chain.set_options(regmode, job.at("pixel_interpolation") == "nearest" ? UPSP_INTERP_NEAREST : UPSP_INTERP_LINEAR,
                      patcher == "polynomial" ? UPSP_PATCH_POLYNOMIAL : UPSP_PATCH_NONE, true);
    if (job.count("filter") && job.at("filter") != "none") {
      const std::string f = job.at("filter");
      if (f != "gaussian" && f != "box") throw std::invalid_argument("job.txt: filter must be none, gaussian or box");
      chain.set_filter(f == "gaussian" ? 1 : 2, geti("filter_size"));
    }
    if (regmode == UPSP_REG_GIVEN)
      for (int c = 0; c < cameras; ++c) {
        auto m6 = read_all<float>(job_dir + "/cam" + std::to_string(c) + ".warp");
        if (m6.size() != (size_t)number_frames * 6)
          throw std::invalid_argument("job directory: cam" + std::to_string(c) + ".warp holds " + std::to_string(m6.size()) +
                                      " floats, expected 6 per frame");
        chain.set_warp_matrices(c, 0, number_frames, m6.data());
      }

    // ---- phase 1: frame loop (the async reader of psp_process.cpp:867-908 is this read loop) ----
    std::cout << "Processing frames" << std::endl;
    std::vector<std::ifstream> vids(cameras);
    size_t max_fb = frame_bytes;
    for (int c = 0; c < cameras; ++c) {
      if (readers[c]) {
        max_fb = std::max(max_fb, readers[c]->frame_bytes());
        if (readers[c]->unpack_lut()) chain.set_unpack_lut(readers[c]->unpack_lut());   // 10-bit cine
        continue;
      }
      vids[c].open(job_dir + "/cam" + std::to_string(c) + ".frames", std::ios::binary);
      if (!vids[c]) throw std::invalid_argument("Cannot open frames of camera " + std::to_string(c));
    }
    std::vector<uint8_t> buf((size_t)chunk * max_fb);
    for (int off = 0; off < number_frames; off += chunk) {
      const int n = std::min(chunk, number_frames - off);
      if (off % 100 == 0) std::cout << "  Rank 0:: processing frame " << off << std::endl;
      for (int c = 0; c < cameras; ++c) {
        if (readers[c]) {     // stored bytes straight to the GPU: decode / table / hot pixels happen there
          readers[c]->read_packed((unsigned)(first_frame + off), (unsigned)n, buf.data());
          chain.push_frames(c, buf.data(), readers[c]->pixel_format(), off, n);
        } else {
          vids[c].read(reinterpret_cast<char*>(buf.data()), (std::streamsize)((size_t)n * frame_bytes));
          if (!vids[c]) throw std::invalid_argument("frame file of camera " + std::to_string(c) + " is too short");
          chain.push_frames(c, buf.data(), p12 ? UPSP_PIX_PACKED12 : UPSP_PIX_U16, off, n);
        }
        chain.wait_pushes();   // buf is reused for the next camera / chunk; the previous chunk's processing keeps running
      }
      chain.process_frames(off, n);
    }
    std::cout << "Global reduction of rms and avg .." << std::endl;
    chain.finish_phase1();
    std::vector<float> sol_avg_final(msize), sol_rms_final(msize), coverage(msize);
    chain.read_phase1_stats(sol_avg_final.data(), sol_rms_final.data(), coverage.data());
    FlatFiles out(out_dir, true);
    out.write_vector("intensity", nullptr, 0);   // created and left empty, as the reference does (its write-behind is disabled, :988)
    out.write_vector("intensity_rms", sol_rms_final.data(), msize);
    out.write_vector("intensity_avg", sol_avg_final.data(), msize);
    out.write_vector("coverage", coverage.data(), msize);

    {
      auto xyz = read_all<float>(job_dir + "/xyz.f32", false);   // psp_process.cpp:1549-1556
      if (xyz.size() == (size_t)msize * 3) {
        std::vector<float> axis(msize);
        const char* names[3] = {"X", "Y", "Z"};
        for (int d = 0; d < 3; ++d) {
          for (int n = 0; n < msize; ++n) axis[n] = xyz[(size_t)n * 3 + d];
          out.write_vector(names[d], axis.data(), msize);
        }
      }
    }

    std::cout << "Construct the transpose" << std::endl;
    chain.global_transpose();
    {
      const size_t rows = std::max<size_t>(1, (256u << 20) / ((size_t)number_frames * 4));
      std::vector<float> blk(rows * number_frames), sol1(msize);
      for (size_t n0 = 0; n0 < (size_t)msize; n0 += rows) {
        const size_t n = std::min(rows, (size_t)msize - n0);
        chain.read_intensity_transpose((int)n0, (int)n, blk.data());
        out.write_block("intensity_transpose", blk.data(), n0, n, number_frames);
        for (size_t r = 0; r < n; ++r) sol1[n0 + r] = blk[r * number_frames];     // frame 1 of every node
      }
      // "a sample Iref/I for frame 1" (psp_process.cpp:1946-1951): avg / sol1 in float, - 1.0 in double
      for (int i = 0; i < msize; ++i) sol1[i] = (float)((double)(sol_avg_final[i] / sol1[i]) - 1.0);
      out.write_vector("intensity_ratio_0", sol1.data(), msize);
      // regression samples (psp_process.cpp:2006-2015)
      write_regression_sample(out_dir + "/vv-int-rms.dat", sol_rms_final.data(), (size_t)msize, 1000);
      write_regression_sample(out_dir + "/vv-int-avg.dat", sol_avg_final.data(), (size_t)msize, 1000);
      write_regression_sample(out_dir + "/vv-int-coverage.dat", coverage.data(), (size_t)msize, 1000);
      write_regression_sample(out_dir + "/vv-int-sample1.dat", sol1.data(), (size_t)msize, 1000);
    }

    // ---- phase 2 ----
    std::cout << "Beginning node processing" << std::endl;
    upsp_phase2_params p{};
    const char* names[6] = {"cal_a", "cal_b", "cal_c", "cal_d", "cal_e", "cal_f"};
    for (int i = 0; i < 6; ++i) p.paint_cal[i] = getf(names[i]);
    p.qbar = getf("qbar");
    p.ps = getf("ps");
    p.degree = geti("degree");
    auto steady = read_all<float>(job_dir + "/steady.f32");
    auto model_temp = read_all<float>(job_dir + "/model_temp.f32");
    if ((int)steady.size() != msize || (int)model_temp.size() != msize)
      throw std::invalid_argument("steady / model_temp inconsistent with msize");
    chain.phase2(p, steady.data(), model_temp.data());
    std::vector<float> rms(msize), avg(msize), gain(msize);
    chain.read_phase2_stats(rms.data(), avg.data(), gain.data());
    std::cout << "Writing pressure rms, average, gain flat files" << std::endl;
    out.write_vector("rms", rms.data(), msize);
    out.write_vector("avg", avg.data(), msize);
    out.write_vector("gain", gain.data(), msize);
    write_regression_sample(out_dir + "/vv-cp-rms.dat", rms.data(), (size_t)msize, 1000);     // psp_process.cpp:2594-2597
    write_regression_sample(out_dir + "/vv-cp-avg.dat", avg.data(), (size_t)msize, 1000);
    for (auto& s : steady)     // psp_process.cpp:2566-2570
      if (s > 3.0f) s = std::numeric_limits<float>::quiet_NaN();
    out.write_vector("steady_state", steady.data(), msize);
    out.write_vector("model_temp", model_temp.data(), msize);
    if (!h5_out.empty()) {
      // the two HDF5 files of psp_process.cpp:2400-2420 / 2535-2604: -h5_out (transposed flag, rms / coverage / steady_state /
      // model_temp) and <add_out_dir>/extras.h5 (+ average); grid, /Condition and code_version in both
      auto xyz = read_all<float>(job_dir + "/xyz.f32", false);
      auto tris = read_all<int32_t>(job_dir + "/tris.i32", false);
      if (xyz.size() != (size_t)msize * 3) throw std::invalid_argument("job directory: xyz.f32 is needed for -h5_out");
      std::vector<float> gx(msize), gy(msize), gz(msize);
      for (int n = 0; n < msize; ++n) gx[n] = xyz[(size_t)n * 3], gy[n] = xyz[(size_t)n * 3 + 1], gz[n] = xyz[(size_t)n * 3 + 2];
      auto gets = [&](const char* k) { return job.count(k) ? job.at(k) : std::string(); };
      auto getd = [&](const char* k) { return job.count(k) ? (float)atof(job.at(k).c_str()) : std::numeric_limits<float>::quiet_NaN(); };
      H5TunnelConditions tc;
      tc.test_id = gets("test_id");
      tc.run = job.count("run") ? geti("run") : 0;
      tc.seq = job.count("sequence") ? geti("sequence") : 0;
      tc.alpha = getd("tc_alpha"); tc.beta = getd("tc_beta"); tc.phi = getd("tc_phi"); tc.mach = getd("tc_mach");
      tc.rey = getd("tc_rey"); tc.ptot = getd("tc_ptot"); tc.qbar = p.qbar; tc.ttot = getd("tc_ttot");
      tc.tcavg = getd("tc_tcavg"); tc.ps = p.ps;
      H5CameraSettings cs;
      cs.framerate = job.count("cam_frame_rate") ? geti("cam_frame_rate") : 0;
      cs.fstop = job.count("cam_fstop") ? getf("cam_fstop") : 0.0f;
      cs.exposure = job.count("cam_exposure") ? getf("cam_exposure") : 0.0f;
      {
        std::istringstream fl(gets("cam_focal_lengths"));
        for (float v; fl >> v;) cs.focal_lengths.push_back(v);
      }
      const bool structured = job.count("structured") && geti("structured") != 0;
      const std::string files[2] = {h5_out, out_dir + "/extras.h5"};
      std::cout << "Initializing output files:" << std::endl;
      for (int k = 0; k < 2; ++k) {
        std::cout << "    " << files[k] << std::endl;
        PSPWriter w(files[k], (size_t)msize, /*transposed=*/k == 0);
        if (structured) {
          w.write_structured_grid(gx, gy, gz, read_all<int32_t>(job_dir + "/grid_sizes.i32"), gets("grid_units"));
        } else {
          const auto comps = read_all<int32_t>(job_dir + "/tri_comps.i32", false);
          std::vector<unsigned> tn(tris.begin(), tris.end());
          std::vector<int> cp(comps.begin(), comps.end());
          cp.resize(tn.size() / 3, 0);
          w.write_unstructured_grid(gx, gy, gz, tn, cp, gets("grid_units"));
        }
        w.write_tunnel_conditions(tc);
        w.write_camera_settings(cs);
        w.write_string_attribute("code_version", gets("code_version"));
        if (k == 1) {
          w.write_new_dataset("rms", rms, "delta Cp");
          w.write_new_dataset("average", avg, "delta Cp");
        } else {
          w.write_new_dataset("rms", rms, "delta Cp");
        }
        w.write_new_dataset("coverage", coverage);
        w.write_new_dataset("steady_state", steady, "Cp");
        w.write_new_dataset("model_temp", model_temp, "F");
        w.close();
      }
    }
    std::cout << "Write pressure transpose ..." << std::endl;
    {
      const size_t rows = std::max<size_t>(1, (256u << 20) / ((size_t)number_frames * 4));
      std::vector<float> blk(rows * number_frames);
      for (size_t n0 = 0; n0 < (size_t)msize; n0 += rows) {
        const size_t n = std::min(rows, (size_t)msize - n0);
        chain.read_pressure_transpose((int)n0, (int)n, blk.data());
        out.write_block("pressure_transpose", blk.data(), n0, n, number_frames);
      }
    }
    std::cerr << "## 'pressure_transpose' written" << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "psp_process_b200: " << e.what() << std::endl;
    return 1;   // psp_process.cpp:1397-1415
  }
  return 0;
}
