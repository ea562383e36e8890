// ORACLE (test infrastructure).  Stand-in for the reference's include/Converter.h, found FIRST on the include path of the
// verbatim matcher build: the real header drags in Eigen and g2o, which are absent; src/ORBmatcher.cc uses only the
// declarations below (src/ORBmatcher.cc:1808, BirdviewMatch's projection branch).  The DEFINITIONS come from the
// reference's own src/Converter.cc, cut out by line range at build time (oracle/gen_ref_parts.py).
#ifndef CONVERTER_H
#define CONVERTER_H
#include <opencv2/core/core.hpp>
namespace ORB_SLAM2 {
class Converter {
public:
    static cv::Point3f BirdPixel2BaseXY(const cv::KeyPoint& kp);
    static cv::Point3f BirdPixel2BaseXY(const cv::Point2f& pt);
    static cv::Point2f BaseXY2BirdPixel(const cv::Point3f& p);
    static cv::Point3f BaseXY2CamXYZ(cv::Point3f p);
};
}  // namespace ORB_SLAM2
#endif
