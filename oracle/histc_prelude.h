/* TEST INFRASTRUCTURE ONLY: what the reference's upsp::intensity_histc template (cpp/lib/image_processing.ipp:10-49) needs in
 * scope when it is compiled on its own (see the _ref/histc.o rule of the Makefile): the standard headers and TWO_POW of
 * cpp/include/utils/general_utils.h:16, and the cv::Mat_ of cv_stub/. */
#ifndef UPSP_ORACLE_HISTC_PRELUDE_H
#define UPSP_ORACLE_HISTC_PRELUDE_H
#include <cmath>
#include <cstdint>
#include <vector>
#include <opencv2/opencv.hpp>
#ifndef TWO_POW
#define TWO_POW(p) (1 << (p))
#endif
#endif
