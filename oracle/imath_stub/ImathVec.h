/* TEST INFRASTRUCTURE ONLY: the part of Imath the reference's ray caster names (cpp/raycast/pspRT.cpp, cpp/include/utils/
 * pspRT.h): a 3-float vector with component-wise arithmetic, a box, a line.  With these the reference's own BVH build /
 * traversal and its watertight ray-triangle test compile from the reference tree into oracle/_ref/ref_probe.  The one Imath
 * ALGORITHM the ray caster calls -- intersects(Box3f, Line3f) for pruning -- is replaced by a conservative test (see
 * ImathBoxAlgo.h here): pruning never changes which triangles a ray hits, only how many are tried. */
#ifndef UPSP_ORACLE_IMATH_VEC_STUB
#define UPSP_ORACLE_IMATH_VEC_STUB
#include <cmath>
#include <limits>
namespace Imath {
template <typename T> class Vec3 {
 public:
  T x, y, z;
  Vec3() : x(0), y(0), z(0) {}
  explicit Vec3(T a) : x(a), y(a), z(a) {}
  Vec3(T a, T b, T c) : x(a), y(b), z(c) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
  Vec3 operator+(const Vec3& v) const { return Vec3(x + v.x, y + v.y, z + v.z); }
  Vec3 operator-(const Vec3& v) const { return Vec3(x - v.x, y - v.y, z - v.z); }
  Vec3 operator-() const { return Vec3(-x, -y, -z); }
  Vec3 operator*(T a) const { return Vec3(x * a, y * a, z * a); }
  Vec3 operator*(const Vec3& v) const { return Vec3(x * v.x, y * v.y, z * v.z); }
  Vec3 operator/(T a) const { return Vec3(x / a, y / a, z / a); }
  Vec3& operator+=(const Vec3& v) { x += v.x; y += v.y; z += v.z; return *this; }
  Vec3& operator-=(const Vec3& v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
  Vec3& operator*=(T a) { x *= a; y *= a; z *= a; return *this; }
  Vec3& operator/=(T a) { x /= a; y /= a; z /= a; return *this; }
  bool operator==(const Vec3& v) const { return x == v.x && y == v.y && z == v.z; }
  bool operator!=(const Vec3& v) const { return !(*this == v); }
  T dot(const Vec3& v) const { return x * v.x + y * v.y + z * v.z; }
  Vec3 cross(const Vec3& v) const { return Vec3(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x); }
  T length2() const { return x * x + y * y + z * z; }
  T length() const {
    const T l2 = length2();
    if (l2 < T(2) * std::numeric_limits<T>::min()) {   // Imath: lengthTiny()
      const T ax = std::fabs(x), ay = std::fabs(y), az = std::fabs(z);
      T m = ax > ay ? ax : ay;
      if (az > m) m = az;
      if (m == T(0)) return T(0);
      const T a = ax / m, b = ay / m, c = az / m;
      return m * std::sqrt(a * a + b * b + c * c);
    }
    return std::sqrt(l2);
  }
  const Vec3& normalize() {
    const T l = length();
    if (l != T(0)) { x /= l; y /= l; z /= l; }
    return *this;
  }
  Vec3 normalized() const { Vec3 v(*this); v.normalize(); return v; }
};
template <typename T> Vec3<T> operator*(T a, const Vec3<T>& v) { return Vec3<T>(a * v.x, a * v.y, a * v.z); }
typedef Vec3<float> V3f;
typedef Vec3<double> V3d;
typedef Vec3<int> V3i;
}  // namespace Imath
#endif
