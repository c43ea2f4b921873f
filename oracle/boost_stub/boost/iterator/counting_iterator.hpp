/* TEST INFRASTRUCTURE ONLY: the one Boost template the reference's cpp/include/utils/clustering.h names
 * (boost::counting_iterator, used by otsu_threshold), so that its find_peaks / first_min_threshold can be compiled from
 * the reference tree into oracle/_ref/ref_probe where Boost does not exist. */
#ifndef UPSP_ORACLE_BOOST_COUNTING_ITERATOR_STUB
#define UPSP_ORACLE_BOOST_COUNTING_ITERATOR_STUB
#include <cstddef>
#include <iterator>
namespace boost {
template <typename T>
class counting_iterator {
 public:
  typedef std::input_iterator_tag iterator_category;
  typedef T value_type;
  typedef std::ptrdiff_t difference_type;
  typedef const T* pointer;
  typedef const T& reference;
  explicit counting_iterator(T v = T()) : v_(v) {}
  reference operator*() const { return v_; }
  counting_iterator& operator++() { ++v_; return *this; }
  counting_iterator operator++(int) { counting_iterator t(*this); ++v_; return t; }
  counting_iterator operator+(std::ptrdiff_t n) const { return counting_iterator((T)(v_ + n)); }
  bool operator==(const counting_iterator& o) const { return v_ == o.v_; }
  bool operator!=(const counting_iterator& o) const { return v_ != o.v_; }
 private:
  T v_;
};
}  // namespace boost
#endif
