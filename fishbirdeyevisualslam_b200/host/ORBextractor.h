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

    // Filled by every operator() call like the reference's ComputePyramid (include/ORBextractor.h:85; read by
    // Frame::ComputeStereoMatches, src/Frame.cc:804,894,911): ROI views (19-px frame around them) over ONE pinned host buffer
    // owned by this extractor, refreshed by a DMA copy of the device pyramid.  A Mat taken from here is valid until the next
    // operator() call of the same extractor (clone() it to keep it).  FBE_IMAGE_PYRAMID=0 in the environment skips the copy
    // for callers that never read the pyramid (this fork's monocular + bird-view tracking) and leaves the Mats empty.
    std::vector<cv::Mat> mvImagePyramid;

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
    bool fill_pyramid_;
    unsigned char* pyr_host_;         // pinned (fbe_host_alloc)
    size_t pyr_host_bytes_;
    int pyr_rows_, pyr_cols_;         // image size the views were laid out for
    void LayoutImagePyramid(int rows, int cols, std::vector<unsigned char*>& dst, std::vector<size_t>& steps);
};

}  // namespace ORB_SLAM2

#endif
