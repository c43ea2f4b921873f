/* TEST INFRASTRUCTURE ONLY: declarations of the OpenCV names that cpp/utils/cv_extras.{h,ipp,cpp} and cpp/lib/patches.ipp mention
 * (with behaviour only for what the checkers run: 2-D point +/-, the Euclidean norm of a 2-D point, a rectangular view of a
 * matrix and its minimum), so
 * that this one reference file can be COMPILED from the reference tree and its upsp::fix_hot_pixels -- which only walks a
 * CV_16U cv::Mat with begin<>() / at<>() (see core.hpp here) -- can be run as a checker (oracle/_ref/ref_probe).  Everything
 * else in that file (text layout, colour maps, sub-matrices) is declared but not defined and never called; the link step
 * leaves those symbols unresolved on purpose (-Wl,--unresolved-symbols=ignore-all). */
#ifndef UPSP_ORACLE_CV_ALL_STUB
#define UPSP_ORACLE_CV_ALL_STUB
#include <cmath>
#include <string>
#include "core.hpp"
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_16UC1 2
#define CV_32FC1 5
#define CV_64F 6
#define CV_8S 1
#define CV_16S 3
#define CV_32S 4
#define CV_MAT_DEPTH_MASK 7
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
namespace cv {
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
  Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
  Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
};
template <typename T> Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_;
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
typedef Point3_<int> Point3i;
/* 3-D point arithmetic as OpenCV's types.hpp defines it (element-wise in T; division by a double is carried out in double and
 * narrowed per element, saturate_cast<T>; the norm accumulates in double): upsp::normal / upsp::area of cpp/lib/models.ipp and
 * the node-normal loop of ref_probe run on these */
template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
  T dot(const Point3_& o) const { return x * o.x + y * o.y + z * o.z; }
  Point3_ cross(const Point3_& o) const { return Point3_(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
  Point3_& operator+=(const Point3_& o) { x += o.x; y += o.y; z += o.z; return *this; }
};
template <typename T> Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> Point3_<T> operator-(const Point3_<T>& a) { return Point3_<T>(-a.x, -a.y, -a.z); }
template <typename T, typename S> Point3_<T> operator*(const Point3_<T>& a, S b) { return Point3_<T>((T)(a.x * b), (T)(a.y * b), (T)(a.z * b)); }
template <typename T, typename S> Point3_<T> operator*(S b, const Point3_<T>& a) { return Point3_<T>((T)(b * a.x), (T)(b * a.y), (T)(b * a.z)); }
template <typename T, typename S> Point3_<T> operator/(const Point3_<T>& a, S b) { return Point3_<T>((T)(a.x / b), (T)(a.y / b), (T)(a.z / b)); }
template <typename T> double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }
template <typename T> double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }   /* as OpenCV */
template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T a, T b, T c, T d) : x(a), y(b), width(c), height(d) {}
  Size size() const { return Size((int)width, (int)height); }
};
typedef Rect_<int> Rect;
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
  double& operator[](int i) { return val[i]; }
};
template <typename T, int N> struct Vec {
  T val[N];
  Vec() {}
  Vec(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
  operator Scalar() const;
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<unsigned char, 3> Vec3b;
template <typename T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, sizeof(T) == 1 ? CV_8U : (sizeof(T) == 2 ? CV_16U : CV_32F)) {}
  Mat_(const Mat& m) : Mat(m) {}
  T& operator()(int r, int c) { return this->template at<T>(r, c); }
  const T* begin() const { return reinterpret_cast<const T*>(this->data); }   /* continuous matrices only (what the probes make) */
  const T* end() const { return reinterpret_cast<const T*>(this->data) + (size_t)this->rows * this->cols; }
  Mat_ operator()(const Rect_<int>& r) const { return Mat_(this->roi(r)); }
};
inline Mat Mat::operator()(const Rect_<int>& r) const { return roi(r); }
inline int Mat::depth() const { return CV_MAT_DEPTH(type_); }   /* single-channel matrices only: type == depth */
enum { FONT_HERSHEY_DUPLEX = 2, FILLED = -1, COLORMAP_JET = 2 };
Size getTextSize(const std::string&, int, double, int, int*);
void rectangle(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0);
void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1, int = 8, bool = false);
/* minimum (and maximum) over a CV_8U / CV_16U / CV_32F matrix or view; the locations are not computed */
inline void minMaxLoc(const Mat& m, double* mn, double* mx = 0, Point* = 0, Point* = 0) {
  double lo = 0, hi = 0;
  bool first = true;
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) {
      const double v = m.type() == CV_8U ? (double)m.at<unsigned char>(r, c) : (m.type() == CV_16U ? (double)m.at<unsigned short>(r, c) : (double)m.at<float>(r, c));
      if (first || v < lo) lo = v;
      if (first || v > hi) hi = v;
      first = false;
    }
  if (mn) *mn = lo;
  if (mx) *mx = hi;
}
Scalar mean(const Mat&);
void applyColorMap(const Mat&, Mat&, int);
void Rodrigues(const Mat&, Mat&);
}  // namespace cv
#endif
