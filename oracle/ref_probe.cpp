// ref_probe -- TEST INFRASTRUCTURE ONLY.  A small main() linked against the reference's OWN sources that compile without
// third-party libraries (cpp/lib/upsp_inputs.cpp, non_cv_upsp.cpp, plot3d.cpp, logging.cpp, cpp/utils/general_utils.cpp,
// file_writers.cpp; `make -C oracle ref`, built only where /root/reference exists, output oracle/_ref/ref_probe).  It calls
// the reference's readers and prints what they parsed in the line format of the product's host/inputs_probe, so that
// tests/test_ref_probe.py can hold the product's restatements against the reference itself.  No reference source is
// copied: this file only calls upsp::FileInputs::Load, upsp::PaintCalibration, upsp::read_tunnel_conditions,
// upsp::read_plot3d_scalar_function_file, upsp::read_plot3d_grid_file / write_plot3d_grid_file, upsp::fwrite and the header
// templates upsp::find_peaks / first_min_threshold (cpp/include/utils/clustering.h, with boost_stub/ for its one Boost include),
// and upsp::MrawReader / PSPVideo / unpack_12bit / unpack_10bit (cpp/lib/MrawReader.cpp, PSPVideo.cpp, with cv_stub/ for the
// zero-filled CV_16U cv::Mat they fill), and upsp::fix_hot_pixels (cpp/utils/cv_extras.cpp; the rest of that file is compiled against
// declarations only, cv_stub/opencv2/opencv.hpp, and left unresolved at link time: it is never called), and the ray caster
// rt::BVH / rt::Triangle::intersect (cpp/raycast/pspRT.cpp, pspRTmem.cpp, with imath_stub/ for the 3-float vector, box and line
// it is written in; the box-line pruning test of Imath is replaced by "visit every node"), and the patch geometry templates
// upsp::cluster_points / PatchClusters constructor / threshold_bounds (cpp/lib/patches.ipp, with eigen_stub/ for the int mask
// matrix they mark pixels in and ref_decls.h for the one name they take from projection.h), and upsp::intensity_histc
// (cpp/lib/image_processing.ipp:10-50, compiled on its own into _ref/histc.o by a pipe from the reference tree, see the Makefile), and upsp::normal / upsp::area of a
// triangle (cpp/lib/models.ipp:135-184, _ref/trigeom.o, same way), and upsp::angle_between (cpp/utils/cv_extras.ipp:67-73) with the
// camera weighters BestView / AverageViews (cpp/lib/projection.ipp:222-268, _ref/weighter.o, same way), and the model-temperature
// lines of the reference's main() (cpp/exec/psp_process.cpp:2287-2310, _ref/modeltemp.o, same way), and the per-node loop of its
// detrend design matrix (cpp/lib/filtering.ipp:20-24, _ref/polymat.o, same way), the double -> float finals of both phases and the
// frame-1 ratio sample (psp_process.cpp:1933-1936, :1947-1949, :2543-2547, _ref/finals.o, same way), P3DModel_::adjust_solution
// (cpp/lib/P3DModel.ipp:146-155, _ref/adjust.o, same way), the per-frame tail of phase 1 (psp_process.cpp:1823-1831, _ref/accum.o, same way) and the per-node loop of its
// phase 2 (:2460-2498, _ref/phase2.o, same way; the Eigen solve inside the detrend fit is the oracle's, loaded with dlopen).
#include <cstdio>
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "ref_decls.h"
#include "MrawReader.h"
#include "PSPVideo.h"
#include "grids.h"
#include "non_cv_upsp.h"
#include "patches.h"
#include "plot3d.h"
#include "upsp_inputs.h"
#include "utils/clustering.h"
#include "utils/cv_extras.h"
#include "utils/pspRT.h"
extern "C" {
#include "utils/pspKdtree.h"
}
#include "utils/file_writers.h"

/* psp_process.cpp:611-624 == upsp_matrix_transpose.cpp:70-93: linked from the latter (compiled with -Dmain=... into _ref) */
void apportion(unsigned long int value, unsigned long int nBins, int* start, int* extent);
/* psp_process.cpp:2287-2310 with the constants of :1096-1098, compiled into _ref/modeltemp.o (see the Makefile) */
void ref_model_temperature(upsp::TunnelConditions& tcond, float* wall_out, float* model_out);
/* psp_process.cpp:2460-2498 compiled into _ref/phase2.o; the fitter it calls is declared in phase2_prelude.h */
#include <dlfcn.h>
#include "phase2_prelude.h"
/* cpp/lib/filtering.ipp:20-24 compiled into _ref/polymat.o (see the Makefile) */
void ref_transpoly_fill(unsigned int n_frames_, unsigned int coeffs_, float* out);
/* psp_process.cpp:1813-1819 and :1823-1831 compiled into _ref/accum.o (see the Makefile) */
void ref_blend_cameras(unsigned int c, std::vector<float>& sol, const std::vector<float>& c_sols);
void ref_phase1_accumulate(unsigned int msize, std::vector<float>& sol, const std::vector<unsigned int>& skipped,
                           std::vector<double>& local_sol_rms, std::vector<double>& local_sol_avg);
/* cpp/lib/P3DModel.ipp:146-155 compiled into _ref/adjust.o (see the Makefile) */
void ref_adjust_solution(const std::map<unsigned int, std::vector<unsigned int>>& overlap_pts_, std::vector<float>& sol);
/* psp_process.cpp:1933-1936, :1947-1949, :2543-2547 compiled into _ref/finals.o (see the Makefile) */
void ref_phase1_finals(unsigned int msize, unsigned long int number_frames, const std::vector<double>& sol_avg_partial,
                       const std::vector<double>& sol_rms_partial, std::vector<float>& sol_avg_final, std::vector<float>& sol_rms_final);
void ref_ratio0(std::vector<float>& sol1, const std::vector<float>& sol_avg_final);
void ref_phase2_finals(unsigned int msize, unsigned long int number_frames, const std::vector<double>& avg, const std::vector<double>& rms,
                       const std::vector<double>& gain, std::vector<float>& avg_final, std::vector<float>& rms_final,
                       std::vector<float>& gain_final);
/* cpp/lib/image_processing.ipp:10-49, instantiated for 16-bit frames in _ref/histc.o (see the Makefile) */
namespace upsp {
template <typename T>
void intensity_histc(const cv::Mat_<T>& img, std::vector<int>& edges, std::vector<int>& counts, unsigned int depth, int bins);
/* cpp/lib/models.ipp:135-184, instantiated for float in _ref/trigeom.o (upsp::Triangle: the reference's data_structs.h) */
/* cpp/lib/projection.ipp:222-268, instantiated for float in _ref/weighter.o (declared as oracle/weighter_prelude.h does) */
template <typename T> struct BestView { std::vector<T> operator()(const std::vector<T>& angles) const; };
template <typename T> struct AverageViews { std::vector<T> operator()(const std::vector<T>& angles) const; };
template <typename FP> cv::Point3_<FP> normal(const Triangle<FP>& tri);
template <typename FP> FP area(const Triangle<FP>& tri);
}

template <typename E>
static std::string str(const E& e) {
  std::ostringstream os;
  os << e;
  return os.str();
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string cmd = argv[1], file = argv[2];
  if (cmd == "apportion" && argc < 4) return 2;
  try {
    if (cmd == "deck") {
      upsp::FileInputs fi;
      if (!fi.Load(file)) return 1;
      if (argc > 3 && std::string(argv[3]) == "check" && !fi.check_all()) return 1;
      if (argc > 4 && std::string(argv[3]) == "write") fi.write_file(argv[4]);
      std::printf("test_id %s\nrun %u\nsequence %u\ntunnel %s\ncameras %u\n", fi.test_id.c_str(), fi.run, fi.sequence, fi.tunnel.c_str(),
                  fi.cameras);
      const char* gt = fi.grid_type == upsp::GridType::P3D ? "p3d" : (fi.grid_type == upsp::GridType::Tri ? "tri" : "none");
      std::printf("sds %s\ngrid %s\ngrid_type %s\nnormals %s\ngrid_units %s\nactive_comps %s\n", fi.sds.c_str(), fi.grid.c_str(), gt,
                  fi.normals.c_str(), fi.grid_units.c_str(), fi.active_comps.c_str());
      for (unsigned c = 0; c < fi.cameras; ++c)
        std::printf("camera %u %s %s %s\n", fi.cam_nums[c], fi.camera_filenames[c].c_str(), fi.targets[c].c_str(), fi.cals[c].c_str());
      std::printf("target_patcher %s\nregistration %s\npixel_interpolation %s\nfilter %s\noverlap %s\n", str(fi.target_patcher).c_str(),
                  str(fi.registration).c_str(), str(fi.pixel_interpolation).c_str(), str(fi.filter).c_str(), str(fi.overlap).c_str());
      std::printf("filter_size %d\noblique_angle %.9g\nnumber_frames %d\nout_dir %s\nout_name %s\n", (int)fi.filter_size,
                  (double)fi.oblique_angle, fi.number_frames, fi.out_dir.c_str(), fi.out_name.c_str());
    } else if (cmd == "paintcal") {
      upsp::PaintCalibration pc(file);
      std::printf("a %.9g\nb %.9g\nc %.9g\nd %.9g\ne %.9g\nf %.9g\n", (double)pc.a, (double)pc.b, (double)pc.c, (double)pc.d, (double)pc.e,
                  (double)pc.f);
      if (argc > 4) std::printf("gain %.9g\n", (double)pc.get_gain((float)atof(argv[3]), (float)atof(argv[4])));
    } else if (cmd == "wtd") {
      const upsp::TunnelConditions tc = upsp::read_tunnel_conditions(file);
      std::fflush(stdout);
      float wall = 0.f, mt = 0.f;
      upsp::TunnelConditions work = tc;    // the reference's lines leave ttot + 459.67 - 459.67 (float) behind; the values as read are printed
      ref_model_temperature(work, &wall, &mt);
      std::printf("alpha %.9g\nbeta %.9g\nphi %.9g\nmach %.9g\nrey %.9g\nptot %.9g\nqbar %.9g\nttot %.9g\nps %.9g\ntcavg %.9g\n",
                  (double)tc.alpha, (double)tc.beta, (double)tc.phi, (double)tc.mach, (double)tc.rey, (double)tc.ptot, (double)tc.qbar,
                  (double)tc.ttot, (double)tc.ps, (double)tc.tcavg);
      std::printf("wall_temp %.9g\nmodel_temp %.9g\n", (double)wall, (double)mt);
    } else if (cmd == "p3dfun") {
      const std::vector<float> sol = upsp::read_plot3d_scalar_function_file(file, argc > 3 ? atoi(argv[3]) : -1);
      std::printf("count %zu\n", sol.size());
      for (float v : sol) std::printf("%.9g\n", (double)v);
    } else if (cmd == "vvdump") {
      if (argc < 5) return 2;
      std::ifstream f(file, std::ios::binary | std::ios::ate);
      std::vector<float> v((size_t)f.tellg() / 4);
      f.seekg(0);
      f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * 4));
      std::printf("written %d\n", upsp::fwrite(argv[3], v, atoi(argv[4])));
    } else if (cmd == "mraw") {      // FILE.mraw [first=1] [count=all] [dump.u16]: upsp::MrawReader + PSPVideo, decoded frames
      upsp::PSPVideo video(std::unique_ptr<upsp::VideoReader>(new upsp::MrawReader(file)));
      const unsigned first = argc > 3 ? (unsigned)atoi(argv[3]) : 1;
      const unsigned count = argc > 4 ? (unsigned)atoi(argv[4]) : video.get_number_frames() - first + 1;
      std::printf("width %u\nheight %u\nbit_depth %u\nnum_frames %u\nframe_rate %u\n", video.get_width(), video.get_height(),
                  video.get_bit_depth(), video.get_number_frames(), video.get_frame_rate());
      FILE* dump = argc > 5 ? std::fopen(argv[5], "wb") : nullptr;
      for (unsigned n = first; n < first + count; ++n) {
        cv::Mat fr = video.get_frame(n);
        if (dump) std::fwrite(fr.data, 2, (size_t)fr.rows * fr.cols, dump);
      }
      if (dump) std::fclose(dump);
    } else if (cmd == "unpack") {    // IN.bin BITS N_PIXELS OUT.u16: upsp::unpack_12bit / unpack_10bit on raw packed bytes
      if (argc < 6) return 2;
      const int bits = atoi(argv[3]);
      const size_t npix = (size_t)atol(argv[4]), nbytes = npix * (size_t)bits / 8;
      std::ifstream f(file, std::ios::binary);
      std::vector<uint8_t> packed(nbytes);
      f.read(reinterpret_cast<char*>(packed.data()), (std::streamsize)nbytes);
      cv::Mat out = cv::Mat::zeros(1, (int)npix, CV_16U);
      if (bits == 12) upsp::unpack_12bit(packed.data(), out, nbytes);
      else upsp::unpack_10bit(packed.data(), out, nbytes);
      FILE* o = std::fopen(argv[5], "wb");
      std::fwrite(out.data, 2, npix, o);
      std::fclose(o);
      std::printf("pixels %zu\n", npix);
    } else if (cmd == "raycast") {   // TRIS.f32 [T][9]  RAYS.f32 [n][6]  OUT.bin: rt::BVH(CreateTriangleMesh(tris), 4).intersect per ray
      if (argc < 5) return 2;
      auto read_f32 = [](const char* path) {
        std::ifstream f(path, std::ios::binary | std::ios::ate);
        std::vector<float> v((size_t)f.tellg() / 4);
        f.seekg(0);
        f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * 4));
        return v;
      };
      const std::vector<float> tris = read_f32(argv[2]), rays = read_f32(argv[3]);
      std::vector<std::shared_ptr<rt::Primitive>> prims = rt::CreateTriangleMesh(tris, 3);      // as createBVH, psp_process.cpp:45-53
      rt::BVH scene(prims, 4);
      FILE* o = std::fopen(argv[4], "wb");
      for (size_t i = 0; i < rays.size() / 6; ++i) {
        const Imath::V3f orig(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), dir(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        rt::Ray ray(orig, dir);
        rt::Hit hitrec;
        const int32_t hit = scene.intersect(ray, &hitrec) ? 1 : 0;
        const int32_t prim = hit ? hitrec.primID : -1;
        const float t = hitrec.t;
        std::fwrite(&hit, 4, 1, o);
        std::fwrite(&t, 4, 1, o);
        std::fwrite(&prim, 4, 1, o);
      }
      std::fclose(o);
      std::printf("rays %zu\ntriangles %zu\n", rays.size() / 6, tris.size() / 9);
    } else if (cmd == "apportion") { // VALUE NBINS: the reference's work split of frames / nodes over ranks
      const unsigned long value = std::stoul(argv[2]), nbins = std::stoul(argv[3]);
      std::vector<int> start(nbins), extent(nbins);
      apportion(value, nbins, start.data(), extent.data());
      std::printf("start");
      for (int v : start) std::printf(" %d", v);
      std::printf("\nextent");
      for (int v : extent) std::printf(" %d", v);
      std::printf("\n");
    } else if (cmd == "patches") {   // targets.txt W H boundary buffer [ref.u16 thresh offset]: InitializeImagePatches' geometry
      if (argc < 7) return 2;        // (psp_process.cpp:2125-2163) with the reference's cluster_points / PatchClusters; output as
      std::ifstream f(file);         // host/patch_geometry_probe prints it
      std::vector<upsp::Target> targs;
      upsp::Target t;
      while (f >> t.uv.x >> t.uv.y >> t.diameter) targs.push_back(t);
      const int W = atoi(argv[3]), H = atoi(argv[4]);
      const unsigned bt = (unsigned)atoi(argv[5]), bf = (unsigned)atoi(argv[6]);
      std::vector<std::vector<upsp::Target>> clusters;
      upsp::cluster_points(targs, clusters, (int)(bt + bf));
      upsp::PatchClusters<float> pc(clusters, cv::Size(W, H), bt, bf);
      if (argc >= 10) {
        cv::Mat_<uint16_t> ref(H, W);
        std::ifstream r(argv[7], std::ios::binary);
        r.read(reinterpret_cast<char*>(ref.data), (std::streamsize)((size_t)W * H * 2));
        pc.threshold_bounds(ref, (unsigned)atoi(argv[8]), (unsigned)atoi(argv[9]));
      }
      for (size_t i = 0; i < clusters.size(); ++i) {
        std::printf("cluster %zu %zu\n", i, clusters[i].size());
        for (size_t j = 0; j < pc.bounds_x[i].size(); ++j) std::printf("b %u %u\n", pc.bounds_x[i][j], pc.bounds_y[i][j]);
        for (size_t j = 0; j < pc.internal_x[i].size(); ++j) std::printf("i %u %u\n", pc.internal_x[i][j], pc.internal_y[i][j]);
      }
    } else if (cmd == "kdtree") {    // PTS.f32 [n][3]  TOL  QUERY.f32 [m][3]  OUT.txt: the reference's kd-tree (pspKdtree.c) as the models use it
      if (argc < 6) return 2;        // line i < n:  "range i: sorted ids within TOL of point i";  line n + q: "nearest q: id"
      auto read_f32 = [](const char* path) {
        std::ifstream f(path, std::ios::binary | std::ios::ate);
        std::vector<float> v((size_t)f.tellg() / 4);
        f.seekg(0);
        f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * 4));
        return v;
      };
      const std::vector<float> pts = read_f32(argv[2]), qs = read_f32(argv[4]);
      const float tol = (float)atof(argv[3]);
      union { void* ptr; uint64_t val; } ud;
      kdtree* root = kd_create(3);
      for (size_t i = 0; i < pts.size() / 3; ++i) {                 // P3DModel.ipp:954-967: float coordinates widened by the call
        ud.val = (uint64_t)i;
        kd_insert3(root, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], ud.ptr);
      }
      FILE* o = std::fopen(argv[5], "w");
      for (size_t i = 0; i < pts.size() / 3; ++i) {
        kdres* res = kd_nearest_range3(root, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], tol);     // :1003
        std::vector<uint64_t> ids;
        while (!kd_res_end(res)) {
          ud.ptr = kd_res_item_data(res);
          ids.push_back(ud.val);
          kd_res_next(res);
        }
        kd_res_free(res);
        std::sort(ids.begin(), ids.end());
        std::fprintf(o, "range %zu:", i);
        for (uint64_t v : ids) std::fprintf(o, " %llu", (unsigned long long)v);
        std::fprintf(o, "\n");
      }
      for (size_t q = 0; q < qs.size() / 3; ++q) {
        const double pos[3] = {qs[3 * q], qs[3 * q + 1], qs[3 * q + 2]};                             // psp_process.cpp:92-98
        kdres* res = kd_nearest(root, pos);
        ud.ptr = kd_res_item_data(res);
        kd_res_free(res);
        std::fprintf(o, "nearest %zu: %llu\n", q, (unsigned long long)ud.val);
      }
      std::fclose(o);
      kd_free(root);
      std::printf("points %zu\nqueries %zu\n", pts.size() / 3, qs.size() / 3);
    } else if (cmd == "hotpix") {    // IN.u16 ROWS COLS OUT.u16 [N_FRAMES]: upsp::fix_hot_pixels, defaults 4064 / 512 / 5
      if (argc < 6) return 2;
      const int rows = atoi(argv[3]), cols = atoi(argv[4]), nf = argc > 6 ? atoi(argv[6]) : 1;
      std::ifstream f(file, std::ios::binary);
      FILE* o = std::fopen(argv[5], "wb");
      for (int k = 0; k < nf; ++k) {
        cv::Mat img = cv::Mat::zeros(rows, cols, CV_16U);
        f.read(reinterpret_cast<char*>(img.data), (std::streamsize)((size_t)rows * cols * 2));
        upsp::fix_hot_pixels(img);
        std::fwrite(img.data, 2, (size_t)rows * cols, o);
      }
      std::fclose(o);
      std::printf("frames %d\n", nf);
    } else if (cmd == "hist") {      // FILE.u16 DEPTH [BINS=256]: upsp::intensity_histc as psp_process.cpp:2157 calls it, then first_min_threshold(5)
      if (argc < 4) return 2;
      std::ifstream f(file, std::ios::binary | std::ios::ate);
      const size_t n = (size_t)f.tellg() / 2;
      f.seekg(0);
      cv::Mat_<uint16_t> img(1, (int)n);
      f.read(reinterpret_cast<char*>(img.data), (std::streamsize)(n * 2));
      std::vector<int> edges, counts;
      upsp::intensity_histc(img, edges, counts, (unsigned)atoi(argv[3]), argc > 4 ? atoi(argv[4]) : 256);
      const unsigned fm = upsp::first_min_threshold(counts, 5);
      std::printf("bin_sz %d\nfirst_min %u\nthreshold %u\nedges", edges[1], fm, (unsigned)(edges[fm] + 5));
      for (int e : edges) std::printf(" %d", e);
      std::printf("\ncounts");
      for (int c : counts) std::printf(" %d", c);
      std::printf("\n");
    } else if (cmd == "trigeom") {   // XYZ.f32 [n][3]  TRIS.i32 [t][3]  OUT.f32: per triangle upsp::normal | upsp::area [t][4], then the node
                                     // normals [n][3] summed as TriModel_::Node::get_normal does (TriModel.ipp:1571-1590: adjacent faces
                                     // in ascending index, normal * area, divided by the double norm unless it is 0)
      if (argc < 5) return 2;
      auto rd = [](const char* p, size_t elem) {
        std::ifstream f(p, std::ios::binary | std::ios::ate);
        std::vector<char> b((size_t)f.tellg() / elem * elem);
        f.seekg(0);
        f.read(b.data(), (std::streamsize)b.size());
        return b;
      };
      const std::vector<char> xb = rd(argv[2], 12), tb = rd(argv[3], 12);
      const float* xyz = reinterpret_cast<const float*>(xb.data());
      const int32_t* tris = reinterpret_cast<const int32_t*>(tb.data());
      const size_t n = xb.size() / 12, nt = tb.size() / 12;
      auto P = [&](int32_t i) { return cv::Point3f(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]); };
      std::vector<float> out(4 * nt + 3 * n);
      std::vector<cv::Point3f> acc(n);
      for (size_t t = 0; t < nt; ++t) {
        const upsp::Triangle<float> tri(P(tris[3 * t]), P(tris[3 * t + 1]), P(tris[3 * t + 2]));
        const cv::Point3f nr = upsp::normal(tri);
        const float ar = upsp::area(tri);
        out[4 * t] = nr.x, out[4 * t + 1] = nr.y, out[4 * t + 2] = nr.z, out[4 * t + 3] = ar;
        for (int k = 0; k < 3; ++k) acc[(size_t)tris[3 * t + k]] += nr * ar;
      }
      for (size_t i = 0; i < n; ++i) {
        cv::Point3f v = acc[i];
        if (cv::norm(v) != 0.0) v = v / cv::norm(v);
        out[4 * nt + 3 * i] = v.x, out[4 * nt + 3 * i + 1] = v.y, out[4 * nt + 3 * i + 2] = v.z;
      }
      FILE* o = std::fopen(argv[4], "wb");
      std::fwrite(out.data(), 4, out.size(), o);
      std::fclose(o);
      std::printf("tris %zu nodes %zu\n", nt, n);
    } else if (cmd == "weights") {   // DIR N_CAMS N_NODES average|best: the files of host/projection_weights_probe; per node seen by >= 2
                                     // cameras: upsp::angle_between(position - centre, normal) (cv_extras.ipp:67-73, narrowed to float
                                     // as ProjWeights::get_angle returns it, projection.ipp:984-991) for the cameras in ascending order,
                                     // the reference's weighter on those angles, every value of the row scaled -> DIR/cam<c>.val.ref
      if (argc < 6) return 2;
      const std::string dir = argv[2];
      const int n_cams = atoi(argv[3]), n = atoi(argv[4]);
      const bool best = std::string(argv[5]) == "best";
      auto rd = [&](const std::string& name) {
        std::ifstream f(dir + "/" + name, std::ios::binary | std::ios::ate);
        std::vector<char> b((size_t)f.tellg());
        f.seekg(0);
        f.read(b.data(), (std::streamsize)b.size());
        return b;
      };
      const std::vector<char> xb = rd("xyz.f32"), nb = rd("nrm.f32"), cb = rd("centers.f64");
      const float* xyz = reinterpret_cast<const float*>(xb.data());
      const float* nrm = reinterpret_cast<const float*>(nb.data());
      const double* cen = reinterpret_cast<const double*>(cb.data());
      std::vector<std::vector<char>> rp(n_cams), vals(n_cams);
      for (int c = 0; c < n_cams; ++c) rp[c] = rd("cam" + std::to_string(c) + ".rowptr"), vals[c] = rd("cam" + std::to_string(c) + ".val");
      for (int i = 0; i < n; ++i) {
        std::vector<int> cs;
        std::vector<float> angs;
        const cv::Point3f pos(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), nn(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
        for (int c = 0; c < n_cams; ++c) {
          const int32_t* r = reinterpret_cast<const int32_t*>(rp[c].data());
          if (r[i + 1] == r[i]) continue;
          cs.push_back(c);
          const cv::Point3f dir3 = pos - cv::Point3f((float)cen[3 * c], (float)cen[3 * c + 1], (float)cen[3 * c + 2]);
          angs.push_back((float)upsp::angle_between(dir3, nn));
        }
        if (cs.size() < 2) continue;
        const std::vector<float> w = best ? upsp::BestView<float>()(angs) : upsp::AverageViews<float>()(angs);
        for (size_t k = 0; k < cs.size(); ++k) {
          const int32_t* r = reinterpret_cast<const int32_t*>(rp[cs[k]].data());
          float* v = reinterpret_cast<float*>(vals[cs[k]].data());
          for (int e = r[i]; e < r[i + 1]; ++e) v[e] *= w[k];
        }
      }
      for (int c = 0; c < n_cams; ++c) {
        FILE* o = std::fopen((dir + "/cam" + std::to_string(c) + ".val.ref").c_str(), "wb");
        std::fwrite(vals[c].data(), 1, vals[c].size(), o);
        std::fclose(o);
      }
      std::printf("cameras %d nodes %d\n", n_cams, n);
    } else if (cmd == "phase2") {    // DIR N F DEGREE ORACLE.so: the reference's per-node loop of phase 2 on DIR/{itrans,avg_final,coverage,steady,
                                     // model_temp}.f32, DIR/paint.cal, DIR/run.wtd -> DIR/ref_{ptrans.f32,rms.f64,avg.f64,gain.f64}
      if (argc < 7) return 2;
      const std::string dir = argv[2];
      const unsigned n = (unsigned)atoi(argv[3]), F = (unsigned)atoi(argv[4]), degree = (unsigned)atoi(argv[5]);
      void* so = dlopen(argv[6], RTLD_NOW);
      if (!so) throw std::runtime_error(dlerror());
      OracleFitter fitter;
      fitter.eval = reinterpret_cast<decltype(fitter.eval)>(dlsym(so, "orc_transpoly_eval_fit"));
      auto build = reinterpret_cast<void (*)(unsigned, unsigned, float*)>(dlsym(so, "orc_transpoly_build"));
      if (!fitter.eval || !build) throw std::runtime_error("oracle symbols missing");
      fitter.n_frames = F, fitter.ncoef = degree + 1;
      fitter.A.resize((size_t)F * (degree + 1));
      build(F, degree, fitter.A.data());
      auto rdf = [&](const std::string& name, size_t count) {
        std::vector<float> v(count);
        std::ifstream f(dir + "/" + name, std::ios::binary);
        f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(count * 4));
        if (!f) throw std::runtime_error("short file " + name);
        return v;
      };
      const std::vector<float> itrans = rdf("itrans.f32", (size_t)n * F), avgf = rdf("avg_final.f32", n), cov = rdf("coverage.f32", n),
                               steady = rdf("steady.f32", n), temp = rdf("model_temp.f32", n);
      upsp::PaintCalibration pcal(dir + "/paint.cal");
      upsp::TunnelConditions tcond = upsp::read_tunnel_conditions(dir + "/run.wtd");
      std::vector<double> rms(n, 0.), avg(n, 0.), gain(n, 0.);
      std::vector<float> ptrans((size_t)n * F, 0.f);
      ref_phase2_nodes(n, 0, F, cov, fitter, rms, avg, gain, tcond, steady, pcal, temp, avgf, itrans.data(), ptrans.data());
      auto wr = [&](const std::string& name, const void* d, size_t bytes) {
        FILE* o = std::fopen((dir + "/" + name).c_str(), "wb");
        std::fwrite(d, 1, bytes, o);
        std::fclose(o);
      };
      wr("ref_ptrans.f32", ptrans.data(), ptrans.size() * 4);
      wr("ref_rms.f64", rms.data(), n * 8);
      wr("ref_avg.f64", avg.data(), n * 8);
      wr("ref_gain.f64", gain.data(), n * 8);
      std::printf("nodes %u frames %u qbar %.9g ps %.9g\n", n, F, (double)tcond.qbar, (double)tcond.ps);
    } else if (cmd == "polymat") {   // OUT.f32 N_FRAMES DEGREE: the design matrix TransPolyFitter's constructor fills, column-major [degree+1][F]
      if (argc < 5) return 2;
      const unsigned F = (unsigned)atoi(argv[3]), nc = (unsigned)atoi(argv[4]) + 1;
      std::vector<float> A((size_t)F * nc, -1.f);
      ref_transpoly_fill(F, nc, A.data());
      FILE* o = std::fopen(argv[2], "wb");
      std::fwrite(A.data(), 4, A.size(), o);
      std::fclose(o);
      std::printf("frames %u coeffs %u\n", F, nc);
    } else if (cmd == "finals") {    // DIR N F: DIR/{sum,sumsq}.f64 + DIR/first.f32 -> phase-1 avg | rms | frame-1 ratio; the same two double
                                     // arrays + DIR/gain.f64 as phase-2 sums -> avg | rms | gain; six float arrays [N] in DIR/ref_finals.f32
      if (argc < 5) return 2;
      const std::string dir = argv[2];
      const unsigned n = (unsigned)atoi(argv[3]);
      const unsigned long F = (unsigned long)atol(argv[4]);
      auto rdd = [&](const std::string& name) {
        std::vector<double> v(n);
        std::ifstream f(dir + "/" + name, std::ios::binary);
        f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * 8));
        if (!f) throw std::runtime_error("short file " + name);
        return v;
      };
      const std::vector<double> sum = rdd("sum.f64"), sumsq = rdd("sumsq.f64"), gain = rdd("gain.f64");
      std::vector<float> sol1(n);
      {
        std::ifstream f(dir + "/first.f32", std::ios::binary);
        f.read(reinterpret_cast<char*>(sol1.data()), (std::streamsize)(n * 4));
      }
      std::vector<float> a1(n, 0.f), r1(n, 0.f), a2(n, 0.f), r2(n, 0.f), g2(n, 0.f);
      ref_phase1_finals(n, F, sum, sumsq, a1, r1);
      ref_ratio0(sol1, a1);
      ref_phase2_finals(n, F, sum, sumsq, gain, a2, r2, g2);
      FILE* o = std::fopen((dir + "/ref_finals.f32").c_str(), "wb");
      for (const std::vector<float>* v : {&a1, &r1, &sol1, &a2, &r2, &g2}) std::fwrite(v->data(), 4, n, o);
      std::fclose(o);
      std::printf("nodes %u frames %lu\n", n, F);
    } else if (cmd == "adjust") {    // PAIRS.i32 [(node, other)] N OUT.f32: P3DModel_::adjust_solution on sol[i] = i, i.e. per node the index its
                                     // value is copied from (the remap psp_setup_b200 writes for upsp_gpu_set_overlap_remap)
      if (argc < 5) return 2;
      std::ifstream f(file, std::ios::binary | std::ios::ate);
      std::vector<int32_t> pairs((size_t)f.tellg() / 4);
      f.seekg(0);
      f.read(reinterpret_cast<char*>(pairs.data()), (std::streamsize)(pairs.size() * 4));
      std::map<unsigned int, std::vector<unsigned int>> overlap;
      for (size_t k = 0; k + 1 < pairs.size(); k += 2) overlap[(unsigned)pairs[k]].push_back((unsigned)pairs[k + 1]);
      const unsigned n = (unsigned)atoi(argv[3]);
      std::vector<float> sol(n);
      for (unsigned i = 0; i < n; ++i) sol[i] = (float)i;
      ref_adjust_solution(overlap, sol);
      FILE* o = std::fopen(argv[4], "wb");
      std::fwrite(sol.data(), 4, n, o);
      std::fclose(o);
      std::printf("nodes %u groups %zu\n", n, overlap.size());
    } else if (cmd == "accum") {     // SOLS.f32 [F][N] N SKIPPED.u32 OUT.f64: the per-frame tail of phase 1 over the blended camera solutions of F
                                     // frames in order -> sum of squares [N] | sum [N] (double); the NaN-marked rows are written back to SOLS
      if (argc < 6) return 2;
      const unsigned n = (unsigned)atoi(argv[3]);
      std::ifstream f(file, std::ios::binary | std::ios::ate);
      const size_t F = (size_t)f.tellg() / 4 / n;
      f.seekg(0);
      std::ifstream sf(argv[4], std::ios::binary | std::ios::ate);
      std::vector<unsigned int> skipped((size_t)sf.tellg() / 4);
      sf.seekg(0);
      sf.read(reinterpret_cast<char*>(skipped.data()), (std::streamsize)(skipped.size() * 4));
      std::vector<double> rms(n, 0.0), avg(n, 0.0);
      std::vector<float> sol(n), marked;
      for (size_t k = 0; k < F; ++k) {
        f.read(reinterpret_cast<char*>(sol.data()), (std::streamsize)(n * 4));
        ref_phase1_accumulate(n, sol, skipped, rms, avg);
        marked.insert(marked.end(), sol.begin(), sol.end());
      }
      FILE* o = std::fopen(argv[5], "wb");
      std::fwrite(rms.data(), 8, n, o);
      std::fwrite(avg.data(), 8, n, o);
      std::fclose(o);
      o = std::fopen((std::string(argv[5]) + ".sol").c_str(), "wb");
      std::fwrite(marked.data(), 4, marked.size(), o);
      std::fclose(o);
      std::printf("frames %zu nodes %u skipped %zu\n", F, n, skipped.size());
    } else if (cmd == "blend") {     // DIR N N_CAMS SKIPPED.u32 OUT.f32: DIR/cam<c>.sols.f32 [F][N] (each camera's project_frame result) summed per
                                     // frame in camera order, then the NaN marks -> the intensity rows [F][N]
      if (argc < 7) return 2;
      const std::string dir = argv[2];
      const unsigned n = (unsigned)atoi(argv[3]), n_cams = (unsigned)atoi(argv[4]);
      std::ifstream sf(argv[5], std::ios::binary | std::ios::ate);
      std::vector<unsigned int> skipped((size_t)sf.tellg() / 4);
      sf.seekg(0);
      sf.read(reinterpret_cast<char*>(skipped.data()), (std::streamsize)(skipped.size() * 4));
      std::vector<std::vector<float>> cams(n_cams);
      for (unsigned c = 0; c < n_cams; ++c) {
        std::ifstream f(dir + "/cam" + std::to_string(c) + ".sols.f32", std::ios::binary | std::ios::ate);
        cams[c].resize((size_t)f.tellg() / 4);
        f.seekg(0);
        f.read(reinterpret_cast<char*>(cams[c].data()), (std::streamsize)(cams[c].size() * 4));
      }
      const size_t F = cams[0].size() / n;
      std::vector<double> rms(n, 0.0), avg(n, 0.0);
      FILE* o = std::fopen(argv[6], "wb");
      for (size_t k = 0; k < F; ++k) {
        std::vector<float> sol;
        for (unsigned c = 0; c < n_cams; ++c) {
          const std::vector<float> c_sols(cams[c].begin() + (std::ptrdiff_t)(k * n), cams[c].begin() + (std::ptrdiff_t)((k + 1) * n));
          ref_blend_cameras(c, sol, c_sols);
        }
        ref_phase1_accumulate(n, sol, skipped, rms, avg);
        std::fwrite(sol.data(), 4, n, o);
      }
      std::fclose(o);
      std::printf("frames %zu nodes %u cameras %u\n", F, n, n_cams);
    } else if (cmd == "peaks") {     // FILE.i32 SEPARATION: upsp::find_peaks on the counts and on 1/counts, first_min_threshold
      if (argc < 4) return 2;
      std::ifstream f(file, std::ios::binary | std::ios::ate);
      std::vector<int> counts((size_t)f.tellg() / 4);
      f.seekg(0);
      f.read(reinterpret_cast<char*>(counts.data()), (std::streamsize)(counts.size() * 4));
      const unsigned sep = (unsigned)atoi(argv[3]);
      std::vector<unsigned int> maxp, minp;
      upsp::find_peaks(counts, maxp, sep);
      std::vector<double> inv(counts.size());
      for (size_t i = 0; i < counts.size(); ++i) inv[i] = 1.0 / counts[i];
      upsp::find_peaks(inv, minp, sep);
      std::printf("max_peaks");
      for (unsigned p : maxp) std::printf(" %u", p);
      std::printf("\nmin_peaks");
      for (unsigned p : minp) std::printf(" %u", p);
      std::printf("\nfirst_min %u\n", upsp::first_min_threshold(counts, sep));
    } else if (cmd == "p3dgrid") {   // FILE sp|dp [OUT]: sizes as host/grid_probe prints them, optional re-write
      const bool dp = argc > 3 && std::string(argv[3]) == "dp";
      auto report = [](const auto& g) {
        std::printf("n_zones %u\nn_points %zu\n", g.num_zones(), (size_t)g.x.size());
        for (unsigned z = 0; z < g.num_zones(); ++z) std::printf("zone %u %u %u %u\n", z, g.grid_size[z][0], g.grid_size[z][1], g.grid_size[z][2]);
      };
      if (dp) {
        upsp::StructuredGrid<double> g;
        upsp::read_plot3d_grid_file(file, g);
        report(g);
        if (argc > 4) upsp::write_plot3d_grid_file(argv[4], g);
      } else {
        upsp::StructuredGrid<float> g;
        upsp::read_plot3d_grid_file(file, g);
        report(g);
        if (argc > 4) upsp::write_plot3d_grid_file(argv[4], g);
      }
    } else {
      return 2;
    }
  } catch (const std::exception& e) {
    std::cerr << "ref_probe: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
