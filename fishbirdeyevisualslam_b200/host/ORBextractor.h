// Drop-in replacement for the reference's include/ORBextractor.h: same namespace, class name, constructor, operator(),
// getters and public mvImagePyramid member, so src/Frame.cc and src/Tracking.cc compile and call it unchanged.
// The work happens in libfbe_b200.so (CUDA, sm_100a) behind include/fbe_cabi.h; this header adds no CUDA dependency.
//   reference interface: include/ORBextractor.h:45-111 ; implementation replaced: src/ORBextractor.cc
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>
#include <opencv2/core/core.hpp>

struct fbe_extractor;

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();

    // Compute the ORB features and descriptors on an image.  Mask is ignored, as in the reference.
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                    cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Filled lazily: the pyramid lives on the device; call SyncImagePyramid() before reading it on the host
    // (only Frame::ComputeStereoMatches does, src/Frame.cc:804,894,911 -- the stereo path).
    std::vector<cv::Mat> mvImagePyramid;
    void SyncImagePyramid();

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

private:
    ORBextractor(const ORBextractor&);
    ORBextractor& operator=(const ORBextractor&);
    fbe_extractor* handle_;
};

}  // namespace ORB_SLAM2

#endif
