// xyz_scalar_to_tbl_b200 -- the flat-file outputs of the frame chain as a Tecplot point table, per zone of the
// structured grid.  Same arguments, messages, exit codes and bytes as the reference's two post-processing tools
// (cpp/exec/xyz_scalar_to_tbl.cpp, cpp/exec/xyz_scalar_to_tbl_delta.cpp); tests/test_tbl_tool.py holds the output byte
// for byte against those tools compiled from the reference tree.
//   xyz_scalar_to_tbl_b200 grid.p3d X Y Z scalar output
//   xyz_scalar_to_tbl_b200 -delta grid.p3d X Y Z scalar1 scalar2     -> ./xyz_scalar_delta.tecplot (scalar1 - scalar2)
// grid.p3d: multi-zone unformatted plot3d, little endian (only its header is read); X / Y / Z / scalar: raw f32 [N]
// (X, Y, Z, steady_state, rms, ... as written by psp_process / psp_process_b200).  NaN scalars are written as 0.
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

static bool read_floats(const std::string& path, std::vector<float>& v) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return false;
  v.resize((size_t)f.tellg() / 4);
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * 4));
  return true;
}

int main(int argc, char** argv) {
  const bool delta = argc > 1 && std::string(argv[1]) == "-delta";
  char** a = argv + (delta ? 1 : 0);
  const int n_args = argc - (delta ? 1 : 0);
  if (n_args < 7) {
    if (delta) std::cerr << "Usage: " << argv[0] << " -delta grid.p3d X Y Z scalar1 scalar2" << std::endl;
    else std::cerr << "Usage: " << argv[0] << " [p3d] [X] [Y] [Z] [scalar] [output]" << std::endl;
    return 1;
  }
  const std::string names[6] = {a[1], a[2], a[3], a[4], a[5], a[6]};
  std::ifstream gfile(names[0], std::ios::binary);
  if (!gfile) {
    std::cout << " Unable to open " << names[0] << " for input, Exiting...\n";
    return 1;
  }
  const int n_flat = delta ? 5 : 4;
  std::vector<float> col[5];
  for (int k = 0; k < n_flat; ++k)
    if (!read_floats(names[k + 1], col[k])) {
      std::cout << " Unable to open " << names[k + 1] << " for input, Exiting...\n";
      return 1;
    }
  const std::string out_name = delta ? "xyz_scalar_delta.tecplot" : names[5];
  FILE* ofile = std::fopen(out_name.c_str(), "w");
  if (!ofile) {
    std::cout << " Unable to open " << (delta ? "xyz_scalar_delta.tecplot" : "xyz_scalar.tecplot") << " for output, Exiting...\n";
    return 1;
  }
  std::fputs("TITLE = \"Surface Cp\"\nVARIABLES = \"x\",\"y\",\"z\",\"Scalar\"\n", ofile);

  // header records: [4][ng][4] [12 ng][(j k l) x ng]...
  int32_t head[4] = {0, 0, 0, 0};
  gfile.read(reinterpret_cast<char*>(head), sizeof head);
  const int ng = head[1];
  std::vector<int32_t> dims((size_t)(ng > 0 ? ng : 0) * 3, 0);
  gfile.read(reinterpret_cast<char*>(dims.data()), (std::streamsize)(dims.size() * 4));
  int tot_nodes = 0;
  for (int ig = 0; ig < ng; ++ig) tot_nodes += dims[(size_t)ig * 3] * dims[(size_t)ig * 3 + 1] * dims[(size_t)ig * 3 + 2];
  for (int k = 0; k < n_flat; ++k)
    if ((int)col[k].size() != tot_nodes) {
      std::cout << names[k + 1] << " nodes(" << col[k].size() << ") != " << names[0] << " nodes(" << tot_nodes << ")\n Exiting...\n";
      std::fclose(ofile);
      return 1;
    }
  size_t i = 0;
  for (int ig = 0; ig < ng; ++ig) {
    const int j = dims[(size_t)ig * 3], k = dims[(size_t)ig * 3 + 1], l = dims[(size_t)ig * 3 + 2];
    std::fprintf(ofile, "ZONE T=\"grid.%i\",F=POINT, I=%i, J=%i\n", ig + 1, j, k);
    for (int n = 0; n < j * k * l; ++n, ++i) {
      float s = col[3][i];
      if (s != s) s = 0.0f;
      if (delta) {
        float s2 = col[4][i];
        if (s2 != s2) s2 = 0.0f;
        s = s - s2;
      }
      std::fprintf(ofile, "%f %f %f %f \n", col[0][i], col[1][i], col[2][i], s);
    }
  }
  std::fclose(ofile);
  return 0;
}
