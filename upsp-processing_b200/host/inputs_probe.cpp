// inputs_probe -- exercises the host-side input readers from the tests (tests/test_run_inputs.py,
// tests/test_p3d_model.py); one sub-command per reader, plain `key value` lines on stdout.
//   inputs_probe deck FILE [check | write OUT]   upsp_inputs.hpp   FileInputs::Load (+ check_all / write_file)
//   inputs_probe paintcal FILE [T Pss]    run_inputs.hpp    PaintCalibration (+ get_gain)
//   inputs_probe wtd FILE                 run_inputs.hpp    read_tunnel_conditions + model_temperature
//   inputs_probe tgts FILE [LABEL]        run_inputs.hpp    read_psp_target_file
//   inputs_probe p3dfun FILE [seps]       run_inputs.hpp    read_plot3d_scalar_function_file
//   inputs_probe hist FILE.u16 DEPTH      run_inputs.hpp    intensity_histc(256 bins) + first_min_threshold(5)
//   inputs_probe overlap GRID TOL [dump]  p3d_model.hpp     P3DModel: overlap groups, src_index, triangles, normals
//   inputs_probe vvdump IN.f32 OUT.dat MAXELS   run_inputs.hpp  write_regression_sample (the vv-*.dat files)
//   inputs_probe interp GRID TOL DATA.f32 XYZ.f32 K OUT.f32   interpolation.hpp   upsp::interpolate onto the points of XYZ
#include <cstdio>
#include <cstring>
#include <iostream>

#include "interpolation.hpp"
#include "p3d_model.hpp"
#include "run_inputs.hpp"
#include "upsp_inputs.hpp"

using namespace upsp_b200;

static void dump(const std::string& path, const void* d, size_t bytes) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write " + path);
  std::fwrite(d, 1, bytes, f);
  std::fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "usage: inputs_probe deck|paintcal|wtd|tgts|p3dfun|hist|overlap FILE ...\n";
    return 1;
  }
  const std::string cmd = argv[1], file = argv[2];
  try {
    if (cmd == "deck") {
      FileInputs fi;
      if (!fi.Load(file)) {
        std::cerr << fi.error << "\n";
        return 1;
      }
      if (argc > 3 && std::string(argv[3]) == "check" && !fi.check_all()) {
        std::cerr << fi.error << "\n";
        return 1;
      }
      if (argc > 4 && std::string(argv[3]) == "write") fi.write_file(argv[4], "1/2/2026");
      std::printf("version %s\ntest_id %s\nrun %d\nsequence %d\ntunnel %s\ncameras %u\n", fi.version.c_str(), fi.test_id.c_str(),
                  fi.run, fi.sequence, fi.tunnel.c_str(), fi.cameras);
      std::printf("sds %s\ngrid %s\ngrid_type %s\nnormals %s\ngrid_units %s\nactive_comps %s\n", fi.sds.c_str(), fi.grid.c_str(),
                  to_string(fi.grid_type), fi.normals.c_str(), fi.grid_units.c_str(), fi.active_comps.c_str());
      for (unsigned c = 0; c < fi.cameras; ++c)
        std::printf("camera %u %s %s %s\n", fi.cam_nums[c], fi.camera_filenames[c].c_str(), fi.targets[c].c_str(), fi.cals[c].c_str());
      std::printf("target_patcher %s\nregistration %s\npixel_interpolation %s\nfilter %s\noverlap %s\n", display_name(fi.target_patcher),
                  display_name(fi.registration), display_name(fi.pixel_interpolation), display_name(fi.filter), display_name(fi.overlap));
      std::printf("filter_size %d\noblique_angle %.9g\nnumber_frames %d\nout_dir %s\nout_name %s\n", fi.filter_size,
                  (double)fi.oblique_angle, fi.number_frames, fi.out_dir.c_str(), fi.out_name.c_str());
    } else if (cmd == "paintcal") {
      PaintCalibration pc(file);
      std::printf("a %.9g\nb %.9g\nc %.9g\nd %.9g\ne %.9g\nf %.9g\n", (double)pc.a, (double)pc.b, (double)pc.c, (double)pc.d, (double)pc.e, (double)pc.f);
      if (argc > 4) std::printf("gain %.9g\n", (double)pc.get_gain((float)atof(argv[3]), (float)atof(argv[4])));
    } else if (cmd == "wtd") {
      const TunnelConditions tc = read_tunnel_conditions(file, &std::cerr);
      float wall = 0.f;
      const float mt = model_temperature(tc, &wall);
      std::printf("alpha %.9g\nbeta %.9g\nphi %.9g\nmach %.9g\nrey %.9g\nptot %.9g\nqbar %.9g\nttot %.9g\nps %.9g\ntcavg %.9g\n",
                  (double)tc.alpha, (double)tc.beta, (double)tc.phi, (double)tc.mach, (double)tc.rey, (double)tc.ptot, (double)tc.qbar,
                  (double)tc.ttot, (double)tc.ps, (double)tc.tcavg);
      std::printf("wall_temp %.9g\nmodel_temp %.9g\n", (double)wall, (double)mt);
    } else if (cmd == "tgts") {
      std::vector<ModelTarget> t;
      if (!read_psp_target_file(file, t, false, argc > 3 ? argv[3] : "*Targets")) {
        std::cerr << "Cannot open '" << file << "'\n";
        return 1;
      }
      std::printf("count %zu\n", t.size());
      for (const auto& q : t) std::printf("target %d %.17g %.17g %.17g %.17g\n", q.num, q.x, q.y, q.z, q.diameter);
    } else if (cmd == "p3dfun") {
      const auto sol = read_plot3d_scalar_function_file(file, argc > 3 ? atoi(argv[3]) : -1);
      std::printf("count %zu\n", sol.size());
      for (float v : sol) std::printf("%.9g\n", (double)v);
    } else if (cmd == "hist") {
      if (argc < 4) throw std::invalid_argument("hist FILE.u16 DEPTH");
      std::ifstream f(file, std::ios::binary | std::ios::ate);
      if (!f) throw std::invalid_argument("Cannot open '" + file + "'");
      std::vector<uint16_t> img((size_t)f.tellg() / 2);
      f.seekg(0);
      f.read(reinterpret_cast<char*>(img.data()), (std::streamsize)(img.size() * 2));
      std::vector<int> edges, counts;
      intensity_histc(img.data(), img.size(), edges, counts, (unsigned)atoi(argv[3]), argc > 4 ? atoi(argv[4]) : 256);
      std::vector<unsigned> peaks;
      find_peaks(counts, peaks, 5);
      const unsigned fm = first_min_threshold(counts, 5);
      std::printf("bin_sz %d\nfirst_min %u\nthreshold %u\npeaks", edges[1], fm, (unsigned)(edges[fm] + 5));
      for (unsigned p : peaks) std::printf(" %u", p);
      std::printf("\nedges");
      for (int e : edges) std::printf(" %d", e);
      std::printf("\ncounts");
      for (int c : counts) std::printf(" %d", c);
      std::printf("\n");
    } else if (cmd == "overlap") {
      if (argc < 4) throw std::invalid_argument("overlap GRID TOL [dump_prefix]");
      const P3DModel model(file, (float)atof(argv[3]));
      int superceded = 0;
      for (int n = 0; n < model.size(); ++n) superceded += model.is_superceded(n);
      std::printf("n_nodes %d\nn_zones %d\nn_faces %d\nn_vert %d\nn_overlapping %zu\nn_superceded %d\n", model.size(), model.num_zones(),
                  model.number_of_faces(), model.number_of_vertices(), model.overlap_pts().size(), superceded);
      for (int z = 0; z < model.num_zones(); ++z)
        std::printf("zone %d %d %d %d\n", z, model.zone_size(z, 0), model.zone_size(z, 1), model.zone_start_idx(z));
      if (argc > 4) {
        const std::string p = argv[4];
        const auto src = model.overlap_src_index();
        std::vector<float> tris;
        std::vector<int> tri_nodes;
        model.extract_tris(tris, tri_nodes);
        std::vector<int32_t> pairs;   // (node, other) for every entry of overlap_pts, map order
        for (const auto& kv : model.overlap_pts())
          for (int o : kv.second) pairs.push_back(kv.first), pairs.push_back(o);
        dump(p + ".src", src.data(), src.size() * 4);
        dump(p + ".trinodes", tri_nodes.data(), tri_nodes.size() * 4);
        dump(p + ".nrm", model.get_n().data(), model.get_n().size() * 4);
        dump(p + ".pairs", pairs.data(), pairs.size() * 4);
      }
    } else if (cmd == "peaks") {     // FILE.i32 SEPARATION
      if (argc < 4) throw std::invalid_argument("peaks FILE.i32 SEPARATION");
      std::ifstream f(file, std::ios::binary | std::ios::ate);
      if (!f) throw std::invalid_argument("Cannot open '" + file + "'");
      std::vector<int> counts((size_t)f.tellg() / 4);
      f.seekg(0);
      f.read(reinterpret_cast<char*>(counts.data()), (std::streamsize)(counts.size() * 4));
      const unsigned sep = (unsigned)atoi(argv[3]);
      std::vector<unsigned> maxp, minp;
      find_peaks(counts, maxp, sep);
      std::vector<double> inv(counts.size());
      for (size_t i = 0; i < counts.size(); ++i) inv[i] = 1.0 / counts[i];
      find_peaks(inv, minp, sep);
      std::printf("max_peaks");
      for (unsigned p : maxp) std::printf(" %u", p);
      std::printf("\nmin_peaks");
      for (unsigned p : minp) std::printf(" %u", p);
      std::printf("\nfirst_min %u\n", first_min_threshold(counts, sep));
    } else if (cmd == "vvdump") {
      if (argc < 5) throw std::invalid_argument("vvdump IN.f32 OUT.dat MAXELS");
      std::ifstream f(file, std::ios::binary | std::ios::ate);
      if (!f) throw std::invalid_argument("Cannot open '" + file + "'");
      std::vector<float> v((size_t)f.tellg() / 4);
      f.seekg(0);
      f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * 4));
      std::printf("written %d\n", write_regression_sample(argv[3], v.data(), v.size(), atoi(argv[4])));
    } else if (cmd == "interp") {
      if (argc < 8) throw std::invalid_argument("interp GRID TOL DATA.f32 XYZ.f32 K OUT.f32");
      auto read_f32 = [](const std::string& path) {
        std::ifstream f(path, std::ios::binary | std::ios::ate);
        if (!f) throw std::invalid_argument("Cannot open '" + path + "'");
        std::vector<float> v((size_t)f.tellg() / 4);
        f.seekg(0);
        f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * 4));
        return v;
      };
      const P3DModel model(file, (float)atof(argv[3]));
      const auto data = read_f32(argv[4]), xyz = read_f32(argv[5]);
      const auto out = interpolate(model, data, xyz.data(), (int)(xyz.size() / 3), (unsigned)atoi(argv[6]), 2.0f);
      dump(argv[7], out.data(), out.size() * 4);
      std::printf("n_out %zu\n", out.size());
    } else {
      std::cerr << "inputs_probe: unknown sub-command '" << cmd << "'\n";
      return 1;
    }
  } catch (const std::exception& e) {
    std::cerr << "inputs_probe: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
