// Drop-in bodies for two Frame methods.
// Frame::AssignFeaturesToGrid (src/Frame.cc:381-411): both grids (64x48 over the undistorted front
// keypoints, 32x32 over the bird keypoints) are bucketed on the device (fbe_grid_assign -> CSR) and copied into the
// reference's mGrid / mGridBirdview vectors, so that GetFeaturesInArea[Birdview] and every other reader see what the
// reference's loop would have produced (same cells, same in-cell order).  Compiled INSIDE the reference tree in place of
// that one method; in this repository it is built against the reference's own Frame.h with the test shim
// oracle/cvshim_m (oracle/Makefile target `dropinmatch`) and checked on the GPU by tests/test_gpu_dropin_match.py.
// PosInGrid / PosInGridBirdview (:548-570) stay as they are: other code calls them per keypoint.
// Frame::UndistortKeyPoints (:638-669): the cv::fisheye::undistortPoints call on the keypoint positions becomes
// fbe_undistort_keypoints (same K, same mDistCoef, P = K); the k1 == 0 copy stays.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "Frame.h"
#include "fbe_cabi.h"
#include "fbe_host.h"

namespace ORB_SLAM2 {

namespace {
static_assert(sizeof(cv::KeyPoint) == sizeof(fbe_keypoint), "cv::KeyPoint must be 28 bytes");

template <int COLS, int ROWS>
void fill_grid(const std::vector<cv::KeyPoint>& keys, int n, float min_x, float min_y, float inv_w, float inv_h,
               std::vector<std::size_t> (&grid)[COLS][ROWS]) {
    std::vector<int> start(COLS * ROWS + 1, 0), items(n > 0 ? n : 1, 0);
    int assigned = 0;
    static thread_local const int dev_rc = fbe_set_device(fbe_host_device());     // same GPU as the extractor and the matchers
    (void)dev_rc;
    if (n > 0 && fbe_grid_assign(reinterpret_cast<const fbe_keypoint*>(keys.data()), n, min_x, min_y, inv_w, inv_h, COLS, ROWS,
                                 start.data(), items.data(), &assigned) != FBE_OK) {
        fprintf(stderr, "Frame::AssignFeaturesToGrid (fbe-b200): %s\n", fbe_last_error());
        abort();      // no CPU fallback
    }
    for (int ix = 0; ix < COLS; ix++)
        for (int iy = 0; iy < ROWS; iy++) {
            const int c = ix * ROWS + iy;
            grid[ix][iy].assign(items.begin() + start[c], items.begin() + start[c + 1]);
        }
}
}  // namespace

void Frame::AssignFeaturesToGrid() {
    fill_grid<FRAME_GRID_COLS, FRAME_GRID_ROWS>(mvKeysUn, N, mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv, mGrid);
    // PosInGridBirdview: round(x * mfGridElementWidthInvBirdview), no origin shift (:560-570)
    fill_grid<FRAME_GRID_BIRD, FRAME_GRID_BIRD>(mvKeysBird, Nbird, 0.f, 0.f, mfGridElementWidthInvBirdview, mfGridElementHeightInvBirdview,
                                                mGridBirdview);
}

void Frame::UndistortKeyPoints() {
    if (mDistCoef.at<float>(0) == 0.0) {
        mvKeysUn = mvKeys;
        return;
    }
    const float K[4] = {mK.at<float>(0, 0), mK.at<float>(1, 1), mK.at<float>(0, 2), mK.at<float>(1, 2)};
    const float D[4] = {mDistCoef.at<float>(0), mDistCoef.at<float>(1), mDistCoef.at<float>(2), mDistCoef.at<float>(3)};
    mvKeysUn.resize(N);
    if (N > 0 && fbe_undistort_keypoints(reinterpret_cast<const fbe_keypoint*>(mvKeys.data()), N, K, D, fbe_host_device(),
                                         reinterpret_cast<fbe_keypoint*>(mvKeysUn.data())) != FBE_OK) {
        fprintf(stderr, "Frame::UndistortKeyPoints (fbe-b200): %s\n", fbe_last_error());
        abort();
    }
}

// Frame::GuidenceKeyBirdPts (:671-684): genEdgesPC() keeps the reference's code (it feeds the ICP); the per-keypoint nearEdges
// filter (:717-739) and the push_back in detection order become one fbe_bird_refine call with no image (filter + ordered
// compaction only).  In the bird-view Frame constructor the following cv::cornerSubPix block (:345-352) folds into the same
// call by passing mBirdviewImg as well -- INTEGRATION.md shows that form.
void Frame::GuidenceKeyBirdPts(std::vector<cv::KeyPoint>& preKeysBird) {
    genEdgesPC();
    const int n = (int)preKeysBird.size();
    if (n == 0) return;
    const cv::Mat& c = mBirdviewContourICP;
    const size_t step = c.rows > 1 ? (size_t)(c.ptr(1) - c.ptr(0)) : (size_t)c.cols;
    std::vector<cv::KeyPoint> kept(n);
    int32_t nkept = 0;
    if (fbe_bird_refine(c.ptr(0), step, NULL, 0, c.rows, c.cols, reinterpret_cast<const fbe_keypoint*>(preKeysBird.data()), n, 5, 5, 40,
                        0.001, fbe_host_device(), NULL, reinterpret_cast<fbe_keypoint*>(kept.data()), &nkept, NULL) != FBE_OK) {
        fprintf(stderr, "Frame::GuidenceKeyBirdPts (fbe-b200): %s\n", fbe_last_error());
        abort();
    }
    mvKeysBird.insert(mvKeysBird.end(), kept.begin(), kept.begin() + nkept);
}

// The reference's bird feature block of the Frame constructor (src/Frame.cc:336-355) -- cv::ORB::create(2000)->detect with
// mBirdviewMask, GuidenceKeyBirdPts' nearEdges filter, cv::cornerSubPix(5x5, 40 iterations, 0.001), ->compute -- as ONE
// device-resident call.  INTEGRATION.md shows the constructor patch: `genEdgesPC();` (it feeds the ICP and used to run inside
// GuidenceKeyBirdPts) followed by this call.  Keypoints, their order, and descriptors are those of OpenCV 4.13's cv::ORB.
void FbeBirdFeatures(const cv::Mat& birdGray, const cv::Mat& birdMask, const cv::Mat& contourICP, std::vector<cv::KeyPoint>& keysBird,
                     cv::Mat& descriptorsBird) {
    struct Handle {
        fbe_bird_orb* h;
        int rows, cols;
        Handle() : h(NULL), rows(0), cols(0) {}
        ~Handle() { if (h) fbe_bird_orb_destroy(h); }
    };
    static thread_local Handle H;
    keysBird.clear();
    descriptorsBird.release();
    if (birdGray.empty()) return;
    if (!H.h || H.rows != birdGray.rows || H.cols != birdGray.cols) {
        if (H.h) fbe_bird_orb_destroy(H.h);
        H.h = NULL;
        if (fbe_bird_orb_create(2000, birdGray.rows, birdGray.cols, 1, fbe_host_device(), &H.h) != FBE_OK) {
            fprintf(stderr, "FbeBirdFeatures (fbe-b200): %s\n", fbe_last_error());
            abort();      // no CPU fallback
        }
        H.rows = birdGray.rows; H.cols = birdGray.cols;
    }
    int32_t cap = 0, n = 0;
    fbe_bird_orb_max_keypoints(H.h, &cap);
    std::vector<cv::KeyPoint> kps(cap);
    std::vector<unsigned char> desc((size_t)cap * 32);
    const size_t istep = birdGray.rows > 1 ? (size_t)(birdGray.ptr(1) - birdGray.ptr(0)) : (size_t)birdGray.cols;
    const size_t mstep = birdMask.rows > 1 ? (size_t)(birdMask.ptr(1) - birdMask.ptr(0)) : (size_t)birdMask.cols;
    const size_t cstep = contourICP.rows > 1 ? (size_t)(contourICP.ptr(1) - contourICP.ptr(0)) : (size_t)contourICP.cols;
    if (fbe_bird_features(H.h, birdGray.ptr(0), istep, 0, birdMask.empty() ? NULL : birdMask.ptr(0), mstep, 0,
                          contourICP.empty() ? NULL : contourICP.ptr(0), cstep, 0, 1, reinterpret_cast<fbe_keypoint*>(kps.data()), &n,
                          desc.data(), NULL) != FBE_OK) {
        fprintf(stderr, "FbeBirdFeatures (fbe-b200): %s\n", fbe_last_error());
        abort();
    }
    kps.resize(n);
    keysBird.swap(kps);
    if (n > 0) {
        descriptorsBird.create(n, 32, CV_8U);
        for (int i = 0; i < n; ++i) std::memcpy(descriptorsBird.ptr(i), &desc[(size_t)i * 32], 32);
    }
}

}  // namespace ORB_SLAM2
