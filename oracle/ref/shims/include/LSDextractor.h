// TEST INFRASTRUCTURE ONLY.  include/peac/AHCPlaneFitter.hpp:39 includes "include/LSDextractor.h" for one type, SurfaceNormal
// (include/LSDextractor.h:34-41; a member vector of it is declared at AHCPlaneFitter.hpp:139 and never used).  This stand-in declares that
// type and keeps the OpenCV line_descriptor module out of the plane-extractor build.
#pragma once
#include <opencv2/opencv.hpp>
#ifdef PSLAM_REF_FULL_HEADERS            // the matcher builds compile the reference's real header (found on the include path after ref/shims)
#include <LSDextractor.h>
#else
class SurfaceNormal {
public:
    cv::Point3f normal;
    cv::Point3f cameraPosition;
    cv::Point2i FramePosition;
    SurfaceNormal() {}
};
#endif
