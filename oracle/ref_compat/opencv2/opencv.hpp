/* oracle/ref_compat: <opencv2/opencv.hpp> (src/Frame.h:29) = everything. */
#ifndef SDORB_REF_COMPAT_OPENCV_HPP
#define SDORB_REF_COMPAT_OPENCV_HPP
#include "core/core.hpp"
#include "features2d/features2d.hpp"
#include "imgproc/imgproc.hpp"
#endif
