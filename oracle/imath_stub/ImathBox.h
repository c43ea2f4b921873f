/* TEST INFRASTRUCTURE ONLY, see ImathVec.h here. */
#ifndef UPSP_ORACLE_IMATH_BOX_STUB
#define UPSP_ORACLE_IMATH_BOX_STUB
#include "ImathVec.h"
namespace Imath {
template <typename V> class Box {
 public:
  V min, max;
  Box() { makeEmpty(); }
  Box(const V& p) : min(p), max(p) {}
  Box(const V& a, const V& b) : min(a), max(b) {}
  void makeEmpty() {
    const float big = std::numeric_limits<float>::max();
    min = V(big, big, big);
    max = V(-big, -big, -big);
  }
  void extendBy(const V& p) {
    for (int i = 0; i < 3; ++i) {
      if (p[i] < min[i]) min[i] = p[i];
      if (p[i] > max[i]) max[i] = p[i];
    }
  }
  void extendBy(const Box& b) {
    for (int i = 0; i < 3; ++i) {
      if (b.min[i] < min[i]) min[i] = b.min[i];
      if (b.max[i] > max[i]) max[i] = b.max[i];
    }
  }
  bool isEmpty() const { return max[0] < min[0] || max[1] < min[1] || max[2] < min[2]; }
  V size() const { return isEmpty() ? V(0, 0, 0) : max - min; }
  V center() const { return (max + min) / 2; }
  bool intersects(const V& p) const {
    for (int i = 0; i < 3; ++i)
      if (p[i] < min[i] || p[i] > max[i]) return false;
    return true;
  }
  unsigned int majorAxis() const {
    unsigned int major = 0;
    const V s = size();
    for (unsigned int i = 1; i < 3; ++i)
      if (s[i] > s[major]) major = i;
    return major;
  }
};
typedef Box<V3f> Box3f;
}  // namespace Imath
#endif
