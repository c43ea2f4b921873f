// psp_setup_b200 -- phase 0 of psp_process for one camera and an unstructured grid, on the GPU:
// grid (.tri) + camera calibration (.json) -> the pixel-to-node projection matrix and the node
// (u, v) file, i.e. InitializeProjection's loop body (cpp/exec/psp_process.cpp:1595-1620) with
// create_projection_mat replaced by upsp_op_create_projection.  The outputs are the inputs of
// psp_process_b200's job directory (cam<c>.rowptr/.col/.val) plus the reference's camNN-uv file.
//
//   psp_setup_b200 -grid FILE.tri -cal FILE.json -out_dir DIR [-oblique_angle 70] [-cam 0] [-device 0]
//   psp_setup_b200 -cal FILE.json -print_cal            (no GPU: prints the parsed calibration)
//
// Deck mode: the whole of psp_process's start-up for a run, from the reference's own command line
// (ParseOpts, cpp/exec/psp_process.cpp:1192-1310) to a complete job directory for psp_process_b200:
//
//   psp_setup_b200 -input_file DECK -paint_cal FILE -job_dir DIR [-steady_p3d FILE] [-model_temp_p3d FILE] [-steady_grid FILE]
//                  [-frames N] [-cutoff_x_max X] [-bound_pts 2] [-buffer_pts 1] [-target_diam_sf 1.2] [-device 0]
//                  [-no_projection]
//
// deck (host/upsp_inputs.hpp) -> video streams and the frame count (InitializeVideoStreams :392-470) -> model
// (unstructured .tri, or structured plot3d through host/p3d_model.hpp: seams -> remap.i32, superceded nodes carry no
// projection row) -> per camera: calibration, projection matrix on the GPU, camNN-uv -> multi-camera weights
// (adjust_projection_for_weights :1631-1641) -> paint calibration, tunnel conditions, model temperature, steady-state
// Cp (phase 2 start-up :2270-2385) -> job.txt, X / Y / Z.  `-no_projection` stops before the GPU step (host-only
// check of everything else).  Function files on an unstructured grid are interpolated from -steady_grid
// (host/interpolation.hpp).  With
// target_patcher = polynomial the visible, projected and sized targets of every camera (getTargets /
// get_target_diameters, host/targets.hpp) are written to DIR/cam<c>.targets.  `@all normals` (structured grids) and
// `@all active_comps` are applied as InitializeModel / phase1 do (psp_process.cpp:2185-2189, 1462-1486).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>

#include "deck_job.hpp"

using namespace upsp_b200;

int main(int argc, char** argv) {
  for (int i = 1; i < argc; ++i)
    if (std::string(argv[i]).rfind("-input_file", 0) == 0) {
      try {
        return run_deck(parse_command_line(argc, argv));
      } catch (const std::exception& e) {
        std::cerr << "psp_setup_b200: " << e.what() << std::endl;
        return 1;
      }
    }
  std::string grid_file, cal_file, out_dir;
  double oblique_angle = 70.0;
  int cam_index = 0, device = 0;
  bool print_cal = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> std::string {
      if (i + 1 >= argc) throw std::invalid_argument("missing value after " + a);
      return argv[++i];
    };
    try {
      if (a == "-grid") grid_file = next();
      else if (a == "-cal") cal_file = next();
      else if (a == "-out_dir") out_dir = next();
      else if (a == "-oblique_angle") oblique_angle = std::atof(next().c_str());
      else if (a == "-cam") cam_index = std::atoi(next().c_str());
      else if (a == "-device") device = std::atoi(next().c_str());
      else if (a == "-print_cal") print_cal = true;
      else throw std::invalid_argument("unknown option " + a);
    } catch (const std::exception& e) {
      std::cerr << "psp_setup_b200: " << e.what() << "\n";
      return 1;
    }
  }
  if (cal_file.empty() || (!print_cal && (grid_file.empty() || out_dir.empty()))) {
    std::cerr << "usage: psp_setup_b200 -grid FILE.tri -cal FILE.json -out_dir DIR [-oblique_angle 70] [-cam 0] [-device 0]\n"
                 "       psp_setup_b200 -cal FILE.json -print_cal\n";
    return 1;
  }
  try {
    const upsp_camera_model cam = read_json_camera_calibration(cal_file);
    std::cout << std::setprecision(17) << "rvec " << cam.rvec[0] << " " << cam.rvec[1] << " " << cam.rvec[2] << "\n"
              << "tvec " << cam.tvec[0] << " " << cam.tvec[1] << " " << cam.tvec[2] << "\n"
              << "K " << cam.fx << " " << cam.fy << " " << cam.cx << " " << cam.cy << "\n"
              << "dist";
    for (double d : cam.dist) std::cout << " " << d;
    std::cout << "\nimageSize " << cam.width << " " << cam.height << std::endl;
    if (print_cal) return 0;
    TriGrid grid = read_tri_grid(grid_file);
    std::cout << "Read " << grid.n_nodes << " nodes and " << grid.n_tris << " faces" << std::endl;
    {   // psp_process loads the grid as TriModel_(file, intersect = true): duplicate nodes collapse (TriModel.ipp:245-255)
      const int initial_size = grid.n_nodes;
      const int overlap = intersect_grid(grid);
      std::cout << "Found " << (initial_size - grid.n_nodes) << " non-unique points\nFound " << overlap << " unique overlapping points" << std::endl;
    }
    std::vector<float> normals;
    calc_normals(grid, normals);
    std::vector<uint8_t> is_data((size_t)grid.n_nodes, 1);
    const float obliqueThresh = (float)((180. - oblique_angle) * 3.14159265358979323846 / 180.0);   // deg2_rad(180. - oblique_angle), :1602
    std::vector<int32_t> code((size_t)grid.n_nodes);
    std::vector<float> uv((size_t)2 * grid.n_nodes);
    if (upsp_op_create_projection(device, &cam, grid.xyz.data(), normals.data(), is_data.data(), grid.n_nodes,
                                  grid.tris.data(), grid.n_tris, obliqueThresh, code.data(), uv.data()) != UPSP_OK)
      throw std::runtime_error(upsp_gpu_last_error());
    std::vector<int32_t> rowptr(1, 0), col;
    for (int n = 0; n < grid.n_nodes; ++n) {
      if (code[(size_t)n] >= 0) col.push_back(code[(size_t)n]);
      rowptr.push_back((int32_t)col.size());
    }
    const std::vector<float> val(col.size(), 1.0f);
    const std::string b = out_dir + "/cam" + std::to_string(cam_index);
    write_all(b + ".rowptr", rowptr);
    write_all(b + ".col", col);
    write_all(b + ".val", val);
    char name[64];
    std::snprintf(name, sizeof name, "/cam%02d-uv", cam_index + 1);                              // FilenameWithCameraPrefix(c + 1, "uv")
    write_all(out_dir + name, uv);
    std::cout << "projected " << grid.n_nodes << " model nodes, accepted " << col.size() << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "psp_setup_b200: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
