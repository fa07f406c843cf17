/* oracle/ref_compat: <opencv/cv.h> (src/ORBextractor.h:30) = the umbrella header of the old C API. */
#ifndef SDORB_REF_COMPAT_CV_H
#define SDORB_REF_COMPAT_CV_H
#include "../opencv2/core/core.hpp"
#include "../opencv2/features2d/features2d.hpp"
#include "../opencv2/imgproc/imgproc.hpp"
#endif
