/* TEST INFRASTRUCTURE ONLY: the two camera-weighting functors whose operator() the reference defines in
 * cpp/lib/projection.ipp:222-268 (declared in cpp/include/projection.h:185-221, a header that needs OpenCV's calibration module,
 * Boost and Eigen's sparse algorithms): angles of one node to the cameras that see it -> weights.  Leaves namespace upsp OPEN for
 * the piped lines (see the _ref/weighter.o rule of the Makefile). */
#include <vector>
namespace upsp {
template <typename T> struct BestView { std::vector<T> operator()(const std::vector<T>& angles) const; };
template <typename T> struct AverageViews { std::vector<T> operator()(const std::vector<T>& angles) const; };
