/* TEST INFRASTRUCTURE ONLY: what the reference's upsp::normal(Triangle) / normal(Polygon) / area(Triangle) templates
 * (cpp/lib/models.ipp:135-184) need in scope when they are compiled on their own (see the _ref/trigeom.o rule of the Makefile):
 * the reference's own aggregates (cpp/include/data_structs.h, against eigen_stub/ and cv_stub/) and the 3-D point of cv_stub/.
 * Leaves namespace upsp OPEN for the piped lines. */
#include <array>
#include <cassert>
#include <cmath>
#include <vector>
#include "ref_decls.h"
#include "data_structs.h"
namespace upsp {
