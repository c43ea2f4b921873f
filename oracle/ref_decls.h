/* TEST INFRASTRUCTURE ONLY: the reference's patches.h includes its projection.h, whose own includes (camera calibration, models,
 * octree, sparse matrices) need OpenCV's calibration module, Boost and Eigen's sparse algorithms.  The patch geometry uses one
 * thing from it, upsp::contains(cv::Size, cv::Point2i) (defined in cpp/lib/projection.cpp, which IS compiled).  So projection.h
 * is skipped through its own include guard and the one declaration is given here. */
#ifndef UPSP_ORACLE_REF_DECLS_H
#define UPSP_ORACLE_REF_DECLS_H
#define UFML_PROJECTION_H_
#include <opencv2/opencv.hpp>
namespace upsp { bool contains(const cv::Size& sz, const cv::Point2i& pt); }
#endif
