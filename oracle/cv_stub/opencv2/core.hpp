/* TEST INFRASTRUCTURE ONLY: the sliver of OpenCV's core module that the reference's camera-file readers name
 * (cpp/lib/PSPVideo.cpp, MrawReader.cpp: a CV_16U cv::Mat that is zero-filled, iterated linearly and handed back by value;
 * cv::Size; CV_Assert), so that the reference's own unpack_12bit / unpack_10bit and .mraw/.cih reader can be compiled from
 * the reference tree into oracle/_ref/ref_probe where OpenCV's C++ headers do not exist.  A Mat here is a continuous
 * row-major buffer shared between copies, as OpenCV's is. */
#ifndef UPSP_ORACLE_CV_CORE_STUB
#define UPSP_ORACLE_CV_CORE_STUB
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <vector>
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); } while (0)
typedef unsigned char uchar;
namespace cv {
struct Scalar;
template <typename T> struct Rect_;
struct Size {
  int width = 0, height = 0;
  Size() = default;
  Size(int w, int h) : width(w), height(h) {}
};
template <typename T>
using MatIterator_ = T*;
class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  Mat() = default;
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type), stride_((size_t)c),
                                buf_(std::make_shared<std::vector<unsigned char>>((size_t)r * c * elem(type), 0)) {
    data = buf_->data();
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  int type() const { return type_; }
  bool empty() const { return data == nullptr; }
  Size size() const { return Size(cols, rows); }
  template <typename T> T* begin() { return reinterpret_cast<T*>(data); }
  template <typename T> T* end() { return reinterpret_cast<T*>(data) + (size_t)rows * cols; }
  template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(data)[(size_t)r * stride_ + c]; }
  template <typename T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data)[(size_t)r * stride_ + c]; }
  /* a view of the rectangle (x, y, width, height) sharing this matrix's pixels, as cv::Mat::operator()(Rect) */
  template <typename R> Mat roi(const R& r) const {
    Mat v(*this);
    v.rows = r.height;
    v.cols = r.width;
    v.data = data + ((size_t)r.y * stride_ + (size_t)r.x) * elem(type_);
    return v;
  }
  /* declared for the compiler only (see opencv.hpp here): never defined, never called */
  Mat(Size, int, const Scalar&);
  Mat(Size, int);
  void convertTo(Mat&, int, double = 1, double = 0) const;
  void copyTo(Mat) const;
  Mat clone() const;
  Mat operator()(const Rect_<int>& r) const;   /* defined in opencv.hpp here, where Rect_ is complete */
  int depth() const;
  int channels() const;
  operator int() const;
 public:
  static size_t elem(int type) { return type == CV_8U ? 1 : (type == CV_16U ? 2 : 4); }
  int type_ = 0;
  size_t stride_ = 0;
  std::shared_ptr<std::vector<unsigned char>> buf_;
};
Mat operator-(const Mat&, double);
Mat operator*(double, const Mat&);
Mat operator*(const Mat&, double);
Mat operator/(const Mat&, double);
Mat operator+(const Mat&, const Mat&);
Mat& operator+=(Mat&, const Mat&);
}  // namespace cv
#endif
