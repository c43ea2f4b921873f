/* TEST INFRASTRUCTURE ONLY, see ImathVec.h here. */
#ifndef UPSP_ORACLE_IMATH_LINE_STUB
#define UPSP_ORACLE_IMATH_LINE_STUB
#include "ImathVec.h"
namespace Imath {
template <typename T> class Line3 {
 public:
  Vec3<T> pos, dir;
  Line3() {}
  Line3(const Vec3<T>& p0, const Vec3<T>& p1) : pos(p0), dir(p1 - p0) { dir.normalize(); }
};
typedef Line3<float> Line3f;
}  // namespace Imath
#endif
