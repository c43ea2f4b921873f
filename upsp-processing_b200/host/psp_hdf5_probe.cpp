// psp_hdf5_probe.cpp -- writes a small PSP HDF5 file with host/psp_hdf5.hpp (test driver of tests/test_psp_hdf5.py).
//   psp_hdf5_probe OUT.h5 unstructured|structured N_NODES
#include <cstdlib>
#include <iostream>

#include "psp_hdf5.hpp"

int main(int argc, char** argv) {
  using namespace upsp_b200;
  if (argc < 4) {
    std::cerr << "usage: psp_hdf5_probe OUT.h5 unstructured|structured N_NODES" << std::endl;
    return 2;
  }
  try {
    const bool structured = std::string(argv[2]) == "structured";
    const size_t n = (size_t)std::atol(argv[3]);
    std::vector<float> x(n), y(n), z(n), rms(n), cov(n), steady(n), temp(n);
    for (size_t i = 0; i < n; ++i) {
      x[i] = 0.5f * (float)i;
      y[i] = 1.0f - (float)i;
      z[i] = (float)(i % 7) * 0.125f;
      rms[i] = 0.001f * (float)i;
      cov[i] = (float)(i % 3);
      steady[i] = -0.25f + 0.01f * (float)i;
      temp[i] = 70.0f;
    }
    PSPWriter w(argv[1], n, /*transposed=*/true);
    if (structured) {
      w.write_structured_grid(x, y, z, {(int)(n / 2), 2, 1}, "in");
    } else {
      std::vector<unsigned> tris;
      std::vector<int> comps;
      for (size_t t = 0; t + 2 < n; ++t) {
        tris.insert(tris.end(), {(unsigned)t, (unsigned)(t + 1), (unsigned)(t + 2)});
        comps.push_back((int)(t % 4));
      }
      w.write_unstructured_grid(x, y, z, tris, comps, "in");
    }
    H5TunnelConditions tc;
    tc.test_id = "t11-0377";
    tc.run = 12;
    tc.seq = 3;
    tc.alpha = 4.25f; tc.beta = -0.5f; tc.phi = 0.0f; tc.mach = 0.85f; tc.rey = 3.0f; tc.ptot = 1500.0f; tc.qbar = 450.0f;
    tc.ttot = 65.0f; tc.tcavg = 71.5f; tc.ps = 900.0f;
    w.write_tunnel_conditions(tc);
    H5CameraSettings cs;
    cs.framerate = 10000;
    cs.fstop = 2.8f;
    cs.exposure = 95.0f;
    cs.focal_lengths = {-3512.25f, -3498.5f};
    w.write_camera_settings(cs);
    w.write_string_attribute("code_version", "probe 1.0");
    w.write_new_dataset("rms", rms, "delta Cp");
    w.write_new_dataset("coverage", cov);
    w.write_new_dataset("steady_state", steady, "Cp");
    w.write_new_dataset("model_temp", temp, "F");
    w.close();
  } catch (const std::exception& e) {
    std::cerr << "psp_hdf5_probe: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
