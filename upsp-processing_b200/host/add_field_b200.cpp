// add_field_b200.cpp -- add a flat [rows x extent] float file to an HDF5 file as a chunked dataset.
//
//   add_field_b200 <hdf5 file> <dataset name> <flat file> <extent>
//
// The reference's `add_field` (cpp/exec/add_field.cpp:22-128; batch use docs/sphinx/quick-start.rst:125-160:
// `add_field $trans_h5_file frames $trans_flat_file $number_frames` puts pressure_transpose into the file psp_process
// wrote): same positional arguments, same checks (four arguments, extent > 0, flat-file size a multiple of one row), same
// dataset (NATIVE_FLOAT, dims {rows, extent}, chunk {1, extent}, fill value 0.0, rows in file order).  The reference
// asserts on bad arguments (abort) and, on an HDF5 error, prints "Cannot open hdf5 file '...' : <detail>" and still
// returns 0; here bad arguments print the failed check and exit with 1, an HDF5-side failure prints the reference's
// message and exits with 1 as well (an exit status of 0 with nothing added is not worth mirroring).
// No HDF5 library: host/h5_append.hpp edits the file directly.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdlib>
#include <iostream>

#include "h5_append.hpp"

int main(int argc, char* argv[]) {
  if (argc != 5) {
    std::cerr << "add_field: expected 4 arguments (hdf5 file, dataset name, flat file, extent), got " << argc - 1 << std::endl;
    return 1;
  }
  const char* hdf5FileName = argv[1];
  const char* dataSetName = argv[2];
  const char* flatFileName = argv[3];
  const long extent = std::atol(argv[4]);
  if (extent <= 0) {
    std::cerr << "add_field: extent must be > 0 (got '" << argv[4] << "')" << std::endl;
    return 1;
  }
  const int fd = ::open(flatFileName, O_RDONLY);
  if (fd < 0) {
    std::cerr << "add_field: cannot open flat file '" << flatFileName << "'" << std::endl;
    return 1;
  }
  struct stat st;
  if (::fstat(fd, &st) != 0) {
    std::cerr << "add_field: cannot stat '" << flatFileName << "'" << std::endl;
    return 1;
  }
  const unsigned long nBytesPerRow = (unsigned long)extent * sizeof(float);
  if ((unsigned long)st.st_size % nBytesPerRow != 0) {
    std::cerr << "add_field: size of '" << flatFileName << "' (" << st.st_size << " bytes) is not a multiple of one row (" << nBytesPerRow
              << " bytes)" << std::endl;
    return 1;
  }
  const unsigned long nRows = (unsigned long)st.st_size / nBytesPerRow;
  int rc = 0;
  try {
    upsp_b200::H5Appender file(hdf5FileName);
    file.add_chunked_float_dataset(dataSetName, fd, nRows, (uint64_t)extent);
  } catch (const std::exception& e) {
    std::cout << "Cannot open hdf5 file '" << hdf5FileName << "' : " << e.what() << std::endl;
    rc = 1;
  }
  ::close(fd);
  return rc;
}
