// Host side of the drop-in ORBextractor: marshals cv::Mat / cv::KeyPoint to the C-ABI and back.
// Error behaviour mirrors the reference (src/ORBextractor.cc:1043-1069): empty image -> return with outputs untouched,
// non-8UC1 image -> assert, zero keypoints -> descriptors released.  A C-ABI failure (no GPU, unsupported geometry)
// is fatal -- there is deliberately no CPU fallback.
#include "ORBextractor.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fbe_cabi.h"
#include "fbe_host.h"

namespace ORB_SLAM2 {

static void die(const char* what, int rc) {
    std::fprintf(stderr, "ORBextractor (fbe-b200): %s failed: %d (%s)\n", what, rc, fbe_last_error());
    std::abort();
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      handle_(NULL), fill_pyramid_(true), pyr_host_(NULL), pyr_host_bytes_(0), pyr_rows_(0), pyr_cols_(0) {
    fbe_extractor_cfg cfg;
    cfg.nfeatures = _nfeatures; cfg.scale_factor = _scaleFactor; cfg.nlevels = _nlevels;
    cfg.ini_th_fast = _iniThFAST; cfg.min_th_fast = _minThFAST; cfg.max_batch = 1; cfg.device = fbe_host_device();
    if (const char* p = std::getenv("FBE_IMAGE_PYRAMID")) fill_pyramid_ = std::atoi(p) != 0;
    int rc = fbe_extractor_create(&cfg, &handle_);
    if (rc != FBE_OK) die("fbe_extractor_create", rc);
    int32_t n = 0;
    const float *s, *is, *s2, *is2;
    fbe_extractor_tables(handle_, &n, &s, &is, &s2, &is2);
    mvScaleFactor.assign(s, s + n);
    mvInvScaleFactor.assign(is, is + n);
    mvLevelSigma2.assign(s2, s2 + n);
    mvInvLevelSigma2.assign(is2, is2 + n);
    const int32_t* per;
    fbe_extractor_features_per_level(handle_, &per);
    mnFeaturesPerLevel.assign(per, per + n);
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() {
    mvImagePyramid.clear();
    if (pyr_host_) fbe_host_free(pyr_host_);
    if (handle_) fbe_extractor_destroy(handle_);
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
    if (_image.empty()) return;
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);

    int32_t cap = 0;
    int rc = fbe_extractor_max_keypoints(handle_, image.rows, image.cols, &cap);
    if (rc != FBE_OK) die("fbe_extractor_max_keypoints", rc);
    // cv::KeyPoint is layout-compatible with fbe_keypoint (7 x 4 bytes)
    static_assert(sizeof(cv::KeyPoint) == sizeof(fbe_keypoint), "cv::KeyPoint layout");
    std::vector<cv::KeyPoint> kps(cap);
    std::vector<unsigned char> desc((size_t)cap * 32);
    int32_t n = 0;
    if (fill_pyramid_) {
        // mvImagePyramid is a side effect of operator() in the reference (ComputePyramid, src/ORBextractor.cc:1052): the level
        // copies ride behind detection and description inside the same call
        std::vector<unsigned char*> dst(nlevels);
        std::vector<size_t> steps(nlevels);
        LayoutImagePyramid(image.rows, image.cols, dst, steps);
        rc = fbe_extract_pyramid(handle_, image.ptr(0), image.rows, image.cols, (size_t)image.step, reinterpret_cast<fbe_keypoint*>(kps.data()),
                                 desc.data(), cap, &n, dst.data(), steps.data());
        if (rc != FBE_OK) die("fbe_extract_pyramid", rc);
    } else {
        rc = fbe_extract(handle_, image.ptr(0), image.rows, image.cols, (size_t)image.step, reinterpret_cast<fbe_keypoint*>(kps.data()),
                         desc.data(), cap, &n);
        if (rc != FBE_OK) die("fbe_extract", rc);
    }

    if (n == 0) {
        _descriptors.release();
    } else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), &desc[(size_t)i * 32], 32);
    }
    kps.resize(n);
    _keypoints.swap(kps);
}

// Pinned host storage for mvImagePyramid of a rows x cols image: every level with its 19-px frame, rows padded to 64 bytes;
// mvImagePyramid[l] is the ROI view inside the frame, like the reference's (src/ORBextractor.cc:1107-1132).  Laid out once per image size.
void ORBextractor::LayoutImagePyramid(int rows, int cols, std::vector<unsigned char*>& dst, std::vector<size_t>& steps) {
    std::vector<int32_t> lr(nlevels), lc(nlevels);
    int rc = fbe_pyramid_geometry(handle_, rows, cols, lr.data(), lc.data());
    if (rc != FBE_OK) die("fbe_pyramid_geometry", rc);
    size_t total = 0;
    for (int l = 0; l < nlevels; ++l) {
        steps[l] = ((size_t)lc[l] + 38 + 63) & ~(size_t)63;
        total += steps[l] * (size_t)(lr[l] + 38);
    }
    const bool relayout = total > pyr_host_bytes_ || rows != pyr_rows_ || cols != pyr_cols_;
    if (relayout) {
        for (int l = 0; l < nlevels; ++l) mvImagePyramid[l] = cv::Mat();
        if (total > pyr_host_bytes_) {
            if (pyr_host_) fbe_host_free(pyr_host_);
            void* p = NULL;
            rc = fbe_host_alloc(&p, total);
            if (rc != FBE_OK) die("fbe_host_alloc", rc);
            pyr_host_ = static_cast<unsigned char*>(p);
            pyr_host_bytes_ = total;
        }
        pyr_rows_ = rows; pyr_cols_ = cols;
    }
    size_t off = 0;
    for (int l = 0; l < nlevels; ++l) {
        dst[l] = pyr_host_ + off;
        off += steps[l] * (size_t)(lr[l] + 38);
        if (relayout) {
            cv::Mat padded(lr[l] + 38, lc[l] + 38, CV_8UC1, dst[l], steps[l]);
            mvImagePyramid[l] = padded(cv::Rect(19, 19, lc[l], lr[l]));      // ROI view with the 19-px frame around it, like the reference
        }
    }
}

}  // namespace ORB_SLAM2
