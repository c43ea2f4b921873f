/* TEST INFRASTRUCTURE ONLY, see ImathVec.h here.  Imath's intersects(box, line, hit point) is an algorithm of a library
 * that is absent from this image and is not restated: the reference uses it only to skip BVH nodes a ray cannot touch.
 * Here every node is visited (returns true), so the traversal tries every triangle; a ray's result -- the hit with the
 * smallest t under `hitrec.t < isect->t` in the reference's own traversal order -- can only differ from a pruned
 * traversal if Imath's test were to reject a box whose triangle the ray does hit. */
#ifndef UPSP_ORACLE_IMATH_BOXALGO_STUB
#define UPSP_ORACLE_IMATH_BOXALGO_STUB
#include "ImathBox.h"
#include "ImathLine.h"
namespace Imath {
template <typename T> bool intersects(const Box<Vec3<T>>& b, const Line3<T>& r, Vec3<T>& ip) {
  (void)b;
  ip = r.pos;
  return true;
}
}  // namespace Imath
#endif
