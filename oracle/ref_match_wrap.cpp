// ORACLE (test infrastructure): C entry points around the reference's OWN ORBmatcher, compiled VERBATIM from
// /root/reference/src/ORBmatcher.cc (plus the Frame / KeyFrame / MapPoint / Converter members it calls, cut out of the
// reference sources at build time by oracle/gen_ref_parts.py) against oracle/cvshim_m.  Recipe: oracle/Makefile ->
// oracle/_ref/libfbe_refmatch.so.  The entry points take the same POD arguments as the restated oracle
// (oracle/match_oracle.cpp: orc_*), build real ORB_SLAM2::Frame / KeyFrame / MapPoint objects from them and call the
// reference method, so that tests can demand  restated oracle == reference's compiled code  on the same inputs.
//
// What this file itself supplies are only members that are NOT on the hot path (constructors, trivial accessors of
// MapPoint / MapPointBird / Map that the reference defines in translation units which cannot be compiled here); the
// camera is the identity (fx = fy = 1, cx = cy = 0, Tcw = I), so a map point at world (u, v, 1) projects to exactly (u, v).
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <chrono>
#include <vector>
#include "ORBVocabulary.h"
#include "opencv2/core/core.hpp"
#define private public
#define protected public
#include "Frame.h"
#include "KeyFrame.h"
#include "Map.h"
#include "MapPoint.h"
#include "MapPointBird.h"
#include "ORBmatcher.h"
#include "Initializer.h"
#include "Converter.h"
#undef private
#undef protected

namespace ORB_SLAM2 {

// ---- members the reference defines elsewhere; minimal equivalents for a test harness ---------------------------------
long unsigned int MapPoint::nNextId = 0;
std::mutex MapPoint::mGlobalMutex;
MapPoint::MapPoint(const cv::Mat& Pos, KeyFrame* pRefKF, Map* pMap)
    : mnFirstKFid(0), mnFirstFrame(0), nObs(0), mnTrackReferenceForFrame(0), mnLastFrameSeen(0), mnBALocalForKF(0), mnFuseCandidateForKF(0),
      mnLoopPointForKF(0), mnCorrectedByKF(0), mnCorrectedReference(0), mnBAGlobalForKF(0), mpRefKF(pRefKF), mnVisible(1), mnFound(1),
      mbBad(false), mpReplaced(static_cast<MapPoint*>(NULL)), mfMinDistance(0), mfMaxDistance(0), mpMap(pMap) {
    Pos.copyTo(mWorldPos);
    mNormalVector = cv::Mat::zeros(3, 1, CV_32F);
    mnId = nNextId++;
}
cv::Mat MapPoint::GetWorldPos() { return mWorldPos.clone(); }
cv::Mat MapPoint::GetNormal() { return mNormalVector.clone(); }
cv::Mat MapPoint::GetDescriptor() { return mDescriptor.clone(); }
bool MapPoint::isBad() { return mbBad; }
int MapPoint::Observations() { return nObs; }
bool MapPoint::IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
void MapPoint::AddObservation(KeyFrame* pKF, size_t idx) { if (!mObservations.count(pKF)) { mObservations[pKF] = idx; nObs++; } }
int MapPoint::GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
static std::vector<std::pair<MapPoint*, MapPoint*> > g_replace_log;      // (this, argument) of every Replace call
void MapPoint::Replace(MapPoint* pMP) { g_replace_log.push_back(std::make_pair(this, pMP)); }
float MapPoint::GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }      // src/MapPoint.cc:373-377
float MapPoint::GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }      // src/MapPoint.cc:379-383

long unsigned int MapPointBird::nNextId = 0;
MapPointBird::MapPointBird(const cv::Mat& Pos, KeyFrame* pRefKF, Map* pMap) : mpRefKF(pRefKF), mpMap(pMap) {
    Pos.copyTo(mWorldPos);
    mnId = nNextId++;
}
cv::Mat MapPointBird::GetWorldPos() { return mWorldPos.clone(); }
cv::Mat MapPointBird::GetDescriptor() { return mDescriptor.clone(); }
void MapPointBird::AddObservation(KeyFrame*, size_t) {}
std::map<KeyFrame*, size_t> MapPointBird::GetObservations() { return mObservations; }
void MapPointBird::ComputeDistinctiveDescriptors() {}
void Map::AddMapPointBird(MapPointBird*) {}

bool KeyFrame::isBad() { return mbBad; }
Initializer::Initializer(const Frame&, float sigma, int iterations) : mSigma(sigma), mSigma2(sigma * sigma), mMaxIterations(iterations) {}
void KeyFrame::SetPose(const cv::Mat& Tcw_) { Tcw_.copyTo(Tcw); }
cv::Mat KeyFrame::GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
cv::Mat KeyFrame::GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
cv::Mat KeyFrame::GetCameraCenter() { return cv::Mat::zeros(3, 1, CV_32F); }
void KeyFrame::AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
std::set<MapPoint*> KeyFrame::GetMapPoints() { return std::set<MapPoint*>(); }

}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

namespace {

struct Kp { float x, y, size, angle, response; int32_t octave, class_id; };
struct FrameView {
    const Kp* kps;
    const uint8_t* desc;
    int32_t n;
    float min_x, min_y, inv_w, inv_h;
    int32_t gcols, grows;
};

std::vector<cv::KeyPoint> to_kps(const Kp* k, int n) {
    std::vector<cv::KeyPoint> v(n);
    static_assert(sizeof(cv::KeyPoint) == sizeof(Kp), "KeyPoint layout");
    if (n) std::memcpy(v.data(), k, (size_t)n * sizeof(Kp));
    return v;
}
cv::Mat to_desc(const uint8_t* d, int n) {
    cv::Mat m(std::max(n, 0), 32, CV_8U);
    for (int i = 0; i < n; ++i) std::memcpy(m.ptr(i), d + (size_t)i * 32, 32);
    return m;
}
cv::Mat desc_row(const uint8_t* d) { return to_desc(d, 1); }
cv::Mat point3(float x, float y, float z) { cv::Mat m(3, 1, CV_32F); m.at<float>(0) = x; m.at<float>(1) = y; m.at<float>(2) = z; return m; }

const float kScale = 1.2f;
const int kLevels = 8;

void set_front_statics(const FrameView& f) {
    Frame::fx = 1.f; Frame::fy = 1.f; Frame::cx = 0.f; Frame::cy = 0.f; Frame::invfx = 1.f; Frame::invfy = 1.f;
    Frame::mnMinX = f.min_x; Frame::mnMinY = f.min_y;
    Frame::mnMaxX = f.min_x + (float)FRAME_GRID_COLS / f.inv_w; Frame::mnMaxY = f.min_y + (float)FRAME_GRID_ROWS / f.inv_h;
    Frame::mfGridElementWidthInv = f.inv_w; Frame::mfGridElementHeightInv = f.inv_h;
}

// a reference Frame holding `f` as its undistorted front keypoints, bucketed by the reference's own AssignFeaturesToGrid
void fill_front(Frame& F, const FrameView& f, const float* scale_factors = nullptr) {
    F.N = f.n; F.Nbird = 0;
    F.mvKeys = to_kps(f.kps, f.n); F.mvKeysUn = F.mvKeys;
    F.mDescriptors = to_desc(f.desc, f.n);
    F.mvpMapPoints.assign(f.n, static_cast<MapPoint*>(NULL));
    F.mvbOutlier.assign(f.n, false);
    F.mnScaleLevels = kLevels; F.mfScaleFactor = kScale; F.mfLogScaleFactor = logf(kScale);
    F.mvScaleFactors.assign(kLevels, 1.f);
    for (int i = 0; i < kLevels; ++i) F.mvScaleFactors[i] = scale_factors ? scale_factors[i] : (i ? F.mvScaleFactors[i - 1] * kScale : 1.f);
    F.mTcw = cv::Mat::eye(4, 4, CV_32F);
    F.mb = 0.f; F.mbf = 0.f; F.mThDepth = 0.f;
    F.mvuRight.assign(f.n, -1.f); F.mvDepth.assign(f.n, -1.f);
    F.mpORBvocabulary = NULL; F.mpORBextractorLeft = NULL; F.mpORBextractorRight = NULL; F.mpReferenceKF = NULL;
    F.mK = cv::Mat::eye(3, 3, CV_32F);
    F.AssignFeaturesToGrid();
}

void fill_bird(Frame& F, const FrameView& f) {
    F.N = 0; F.Nbird = f.n;
    F.mvKeysBird = to_kps(f.kps, f.n);
    F.mDescriptorsBird = to_desc(f.desc, f.n);
    F.mvpMapPointsBird.assign(f.n, static_cast<MapPointBird*>(NULL));
    Frame::mfGridElementWidthInvBirdview = f.inv_w; Frame::mfGridElementHeightInvBirdview = f.inv_h;
    Frame::birdviewCols = (int)lroundf((float)FRAME_GRID_BIRD / f.inv_w); Frame::birdviewRows = (int)lroundf((float)FRAME_GRID_BIRD / f.inv_h);
    F.mTcw = cv::Mat::eye(4, 4, CV_32F);
    Frame::Tbc = cv::Mat::eye(4, 4, CV_32F);
    F.AssignFeaturesToGrid();
}

MapPoint* make_mp(float u, float v, const uint8_t* desc, int level, int nobs) {
    MapPoint* p = new MapPoint(point3(u, v, 1.f), NULL, NULL);
    p->mDescriptor = desc_row(desc);
    p->nObs = nobs;
    // PredictScale = ceil(log(mfMaxDistance / dist) / log(1.2)): place the ratio half a level below `level`
    const float dist = sqrtf(u * u + v * v + 1.f);
    p->mfMaxDistance = dist * powf(kScale, (float)level - 0.5f);
    p->mfMinDistance = 0.f;
    p->mNormalVector = point3(u / dist, v / dist, 1.f / dist);        // viewing angle 0: passes the 60 degree test
    return p;
}

struct Pool {       // owns the objects of one call
    std::vector<MapPoint*> mps; std::vector<MapPointBird*> mpbs; std::vector<KeyFrame*> kfs;
    ~Pool() { for (auto p : mps) delete p; for (auto p : mpbs) delete p; for (auto p : kfs) delete p; }
};

bool check_grid(const FrameView& f, int cols, int rows) { return f.gcols == cols && f.grows == rows; }

}  // namespace

static double g_last_method_us = 0.0;

#ifdef FBE_DROPIN
namespace ORB_SLAM2 {
void FbeBirdFeatures(const cv::Mat& birdGray, const cv::Mat& birdMask, const cv::Mat& contourICP, std::vector<cv::KeyPoint>& keysBird,
                     cv::Mat& descriptorsBird);
}
#endif

extern "C" {

// Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:381-411, 548-558) as CSR [ix*rows+iy]
int refm_grid_assign(const Kp* kps, int n, float min_x, float min_y, float inv_w, float inv_h, int gcols, int grows,
                     int32_t* cell_start, int32_t* cell_items) {
    FrameView f{kps, nullptr, n, min_x, min_y, inv_w, inv_h, gcols, grows};
    std::vector<uint8_t> zero((size_t)std::max(n, 1) * 32, 0);
    f.desc = zero.data();
    int off = 0;
    Frame F;
    if (gcols == FRAME_GRID_COLS && grows == FRAME_GRID_ROWS) {
        set_front_statics(f); fill_front(F, f);
        for (int ix = 0; ix < gcols; ++ix)
            for (int iy = 0; iy < grows; ++iy) {
                cell_start[ix * grows + iy] = off;
                for (size_t j = 0; j < F.mGrid[ix][iy].size(); ++j) cell_items[off++] = (int)F.mGrid[ix][iy][j];
            }
    } else if (gcols == FRAME_GRID_BIRD && grows == FRAME_GRID_BIRD) {
        fill_bird(F, f);
        for (int ix = 0; ix < gcols; ++ix)
            for (int iy = 0; iy < grows; ++iy) {
                cell_start[ix * grows + iy] = off;
                for (size_t j = 0; j < F.mGridBirdview[ix][iy].size(); ++j) cell_items[off++] = (int)F.mGridBirdview[ix][iy][j];
            }
    } else return -1;
    cell_start[gcols * grows] = off;
    return off;
}

// Frame::GetFeaturesInArea / GetFeaturesInAreaBirdview (src/Frame.cc:493-546, 572-626)
int refm_features_in_area(const FrameView* f, float x, float y, float r, int minLevel, int maxLevel, int upper_inclusive,
                          int32_t* out, int cap) {
    Frame F;
    std::vector<size_t> v;
    if (upper_inclusive) { if (!check_grid(*f, FRAME_GRID_COLS, FRAME_GRID_ROWS)) return -1; set_front_statics(*f); fill_front(F, *f); v = F.GetFeaturesInArea(x, y, r, minLevel, maxLevel); }
    else { if (!check_grid(*f, FRAME_GRID_BIRD, FRAME_GRID_BIRD)) return -1; fill_bird(F, *f); v = F.GetFeaturesInAreaBirdview(x, y, r, minLevel, maxLevel); }
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:406-521)
int refm_search_for_initialization(const FrameView* f1, const FrameView* f2, float* prev_matched, int32_t* matches12, int window,
                                   float nn_ratio, int check_ori) {
    if (!check_grid(*f2, FRAME_GRID_COLS, FRAME_GRID_ROWS)) return -1;
    set_front_statics(*f2);
    Frame F1, F2;
    fill_front(F1, *f1); fill_front(F2, *f2);
    std::vector<cv::Point2f> prev(f1->n);
    for (int i = 0; i < f1->n; ++i) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher m(nn_ratio, check_ori != 0);
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const int n = m.SearchForInitialization(F1, F2, prev, m12, window);
    g_last_method_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    for (int i = 0; i < f1->n; ++i) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
    return n;
}

// wall time of the ORBmatcher METHOD CALL inside the last refm_search_for_initialization (the harness's Frame construction
// from POD arrays is outside it): used to time config C1 through the C++ class (tools/c1_cpp.py)
double refm_last_method_us() { return g_last_method_us; }

// ORBmatcher::BirdviewMatch, isProject == 0 (src/ORBmatcher.cc:1602-1760)
int refm_birdview_match(const Kp* ref_kps, const uint8_t* ref_desc, int n_ref, const FrameView* cur, int window, float nn_ratio,
                        int check_ori, int32_t* dmatches, int32_t* n_dmatches) {
    if (!check_grid(*cur, FRAME_GRID_BIRD, FRAME_GRID_BIRD)) return -1;
    Frame F;
    fill_bird(F, *cur);
    std::vector<cv::KeyPoint> rk = to_kps(ref_kps, n_ref);
    cv::Mat rd = to_desc(ref_desc, n_ref);
    std::vector<MapPointBird*> rmp(n_ref, static_cast<MapPointBird*>(NULL));
    std::vector<cv::DMatch> dm;
    ORBmatcher m(nn_ratio, check_ori != 0);
    const int n = m.BirdviewMatch(F, rk, rd, rmp, dm, 0, window);
    for (size_t i = 0; i < dm.size(); ++i) { dmatches[3 * i] = dm[i].queryIdx; dmatches[3 * i + 1] = dm[i].trainIdx; dmatches[3 * i + 2] = (int)dm[i].distance; }
    *n_dmatches = (int)dm.size();
    return n;
}

// ORBmatcher::BirdMapPointMatch first pass (src/ORBmatcher.cc:1763-1863).  mp_base: n x 3 positions in the BASE frame
// (Tbw = I here), NaN x = NULL map point.  The reference itself applies the |z| > 0.2 test, Converter::BaseXY2BirdPixel
// and the image test; pix_out returns the pixel it searched around (NaN x = rejected) so that the restated oracle can be
// given exactly the same input.  matches12 is recovered from mvpMapPointsBird with filterSize = +inf and camera points
// equal to the map points... it is returned through the inlier assignment: assigned[k] = map point index or -1.
int refm_bird_map_point_match(const float* mp_base, const uint8_t* mp_desc, int n_mp, const FrameView* cur, int window, float nn_ratio,
                              float* pix_out, int32_t* assigned) {
    if (!check_grid(*cur, FRAME_GRID_BIRD, FRAME_GRID_BIRD)) return -1;
    Frame F;
    fill_bird(F, *cur);
    F.mvKeysBirdCamXYZ.assign(cur->n, cv::Point3f(0.f, 0.f, 0.f));
    Pool pool;
    std::vector<MapPointBird*> mps(n_mp, static_cast<MapPointBird*>(NULL));
    std::map<MapPointBird*, int> index;
    for (int i = 0; i < n_mp; ++i) {
        pix_out[2 * i] = pix_out[2 * i + 1] = std::numeric_limits<float>::quiet_NaN();
        if (std::isnan(mp_base[3 * i])) continue;
        MapPointBird* p = new MapPointBird(point3(mp_base[3 * i], mp_base[3 * i + 1], mp_base[3 * i + 2]), NULL, NULL);
        p->mDescriptor = desc_row(mp_desc + (size_t)i * 32);
        pool.mpbs.push_back(p); mps[i] = p; index[p] = i;
        if (fabs(mp_base[3 * i + 2]) > 0.2) continue;
        const cv::Point2f pt = Converter::BaseXY2BirdPixel(cv::Point3f(mp_base[3 * i], mp_base[3 * i + 1], mp_base[3 * i + 2]));
        if (pt.x < 0 || pt.x >= Frame::birdviewCols || pt.y < 0 || pt.y >= Frame::birdviewRows) continue;
        pix_out[2 * i] = pt.x; pix_out[2 * i + 1] = pt.y;
    }
    ORBmatcher m(nn_ratio, true);
    const int inl = m.BirdMapPointMatch(F, mps, window, std::numeric_limits<float>::infinity());
    for (int k = 0; k < cur->n; ++k) assigned[k] = F.mvpMapPointsBird[k] ? index[F.mvpMapPointsBird[k]] : -1;
    return inl;
}

// ORBmatcher::SearchByProjection(Frame&, const Frame& LastFrame, th, bMono = true) (src/ORBmatcher.cc:1329-1471)
int refm_search_by_projection_last(const FrameView* cur, const Kp* last_kps, const float* last_proj, const uint8_t* last_mp_desc, int n_last,
                                   const float* scale_factors, const uint8_t* cur_taken, const uint8_t* last_has_obs, float th, int check_ori,
                                   int32_t* cur_mp) {
    if (!check_grid(*cur, FRAME_GRID_COLS, FRAME_GRID_ROWS)) return -1;
    set_front_statics(*cur);
    Frame C, L;
    fill_front(C, *cur, scale_factors);
    FrameView lv = *cur; lv.kps = last_kps; lv.n = n_last;
    std::vector<uint8_t> zero((size_t)std::max(n_last, 1) * 32, 0);
    lv.desc = zero.data();
    fill_front(L, lv, scale_factors);
    Pool pool;
    std::map<MapPoint*, int> index;
    for (int i = 0; i < n_last; ++i) {
        if (std::isnan(last_proj[2 * i])) continue;
        MapPoint* p = make_mp(last_proj[2 * i], last_proj[2 * i + 1], last_mp_desc + (size_t)i * 32, 0, (!last_has_obs || last_has_obs[i]) ? 1 : 0);
        pool.mps.push_back(p); L.mvpMapPoints[i] = p; index[p] = i;
    }
    // `blocker` stands for "a map point with observations" on taken keypoints; every other keypoint starts with `weak`, a
    // map point WITHOUT observations (it does not block, :1404-1406), so that afterwards "untouched" (still weak) can be told
    // from "assigned by this call and then set to NULL by the orientation check" (-2 in the C-ABI convention)
    MapPoint* blocker = make_mp(0, 0, zero.data(), 0, 1);
    MapPoint* weak = make_mp(0, 0, zero.data(), 0, 0);
    pool.mps.push_back(blocker); pool.mps.push_back(weak);
    for (int k = 0; k < cur->n; ++k) C.mvpMapPoints[k] = (cur_taken && cur_taken[k]) ? blocker : weak;
    ORBmatcher m(0.9f, check_ori != 0);
    const int n = m.SearchByProjection(C, L, th, true);
    for (int k = 0; k < cur->n; ++k) {
        MapPoint* p = C.mvpMapPoints[k];
        cur_mp[k] = (p == blocker || p == weak) ? -1 : (p ? index[p] : -2);
    }
    return n;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:46-130)
int refm_search_by_projection_map(const FrameView* cur, const float* scale_factors, const float* mp_proj, const int32_t* mp_level,
                                  const float* mp_viewcos, const uint8_t* mp_desc, int n_mp, const uint8_t* cur_taken,
                                  const uint8_t* mp_has_obs, float th, float nn_ratio, int32_t* cur_mp) {
    if (!check_grid(*cur, FRAME_GRID_COLS, FRAME_GRID_ROWS)) return -1;
    set_front_statics(*cur);
    Frame C;
    fill_front(C, *cur, scale_factors);
    Pool pool;
    std::map<MapPoint*, int> index;
    std::vector<MapPoint*> mps(n_mp);
    for (int i = 0; i < n_mp; ++i) {
        MapPoint* p = make_mp(0, 0, mp_desc + (size_t)i * 32, 0, (!mp_has_obs || mp_has_obs[i]) ? 1 : 0);
        p->mbTrackInView = true; p->mTrackProjX = mp_proj[2 * i]; p->mTrackProjY = mp_proj[2 * i + 1];
        p->mnTrackScaleLevel = mp_level[i]; p->mTrackViewCos = mp_viewcos[i];
        pool.mps.push_back(p); mps[i] = p; index[p] = i;
    }
    std::vector<uint8_t> zero(32, 0);
    MapPoint* blocker = make_mp(0, 0, zero.data(), 0, 1);
    pool.mps.push_back(blocker);
    for (int k = 0; k < cur->n; ++k) if (cur_taken && cur_taken[k]) C.mvpMapPoints[k] = blocker;
    ORBmatcher m(nn_ratio, true);
    const int n = m.SearchByProjection(C, mps, th);
    for (int k = 0; k < cur->n; ++k) {
        MapPoint* p = C.mvpMapPoints[k];
        cur_mp[k] = (!p || p == blocker) ? -1 : index[p];
    }
    return n;
}

// relocalisation (src/ORBmatcher.cc:1473-1600, level_up = 1) and loop closing (:291-404, level_up = 0) SearchByProjection
int refm_search_by_projection_kf(const FrameView* cur, const Kp* q_kps, const float* proj, const int32_t* level, const uint8_t* mp_desc,
                                 int n_mp, const float* scale_factors, const uint8_t* cur_taken, float th, int th_dist, int level_up,
                                 int check_ori, int32_t* cur_mp) {
    if (!check_grid(*cur, FRAME_GRID_COLS, FRAME_GRID_ROWS)) return -1;
    set_front_statics(*cur);
    Pool pool;
    std::map<MapPoint*, int> index;
    std::vector<MapPoint*> mps(n_mp, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < n_mp; ++i) {
        if (std::isnan(proj[2 * i])) continue;
        MapPoint* p = make_mp(proj[2 * i], proj[2 * i + 1], mp_desc + (size_t)i * 32, level[i], 1);
        pool.mps.push_back(p); mps[i] = p; index[p] = i;
    }
    std::vector<uint8_t> zero((size_t)std::max(n_mp, 1) * 32, 0);
    MapPoint* blocker = make_mp(0, 0, zero.data(), 0, 1);
    pool.mps.push_back(blocker);
    int n = 0;
    if (level_up == 1) {
        // relocalisation: the map points hang on a key frame whose keypoints supply the angles
        Frame C, K;
        fill_front(C, *cur, scale_factors);
        FrameView kv = *cur; kv.kps = q_kps; kv.n = n_mp; kv.desc = zero.data();
        fill_front(K, kv, scale_factors);
        K.mvpMapPoints = mps;
        KeyFrame* kf = new KeyFrame(K, NULL, NULL);
        pool.kfs.push_back(kf);
        for (int k = 0; k < cur->n; ++k) if (cur_taken && cur_taken[k]) C.mvpMapPoints[k] = blocker;
        ORBmatcher m(0.9f, check_ori != 0);
        std::set<MapPoint*> found;
        n = m.SearchByProjection(C, kf, found, th, th_dist);
        for (int k = 0; k < cur->n; ++k) {
            MapPoint* p = C.mvpMapPoints[k];
            if (cur_taken && cur_taken[k]) cur_mp[k] = p == blocker ? -1 : (p ? index[p] : -2);
            else cur_mp[k] = p ? index[p] : -1;
        }
    } else {
        // loop closing: the searched frame is a key frame; Scw = identity; candidate points come as a list
        Frame K;
        fill_front(K, *cur, scale_factors);
        KeyFrame* kf = new KeyFrame(K, NULL, NULL);
        pool.kfs.push_back(kf);
        std::vector<MapPoint*> pts, matched(cur->n, static_cast<MapPoint*>(NULL));
        std::vector<int> pt_index;
        for (int i = 0; i < n_mp; ++i) if (mps[i]) { pts.push_back(mps[i]); pt_index.push_back(i); }
        for (int k = 0; k < cur->n; ++k) if (cur_taken && cur_taken[k]) matched[k] = blocker;
        ORBmatcher m(0.75f, true);
        n = m.SearchByProjection(kf, cv::Mat::eye(4, 4, CV_32F), pts, matched, (int)th);
        for (int k = 0; k < cur->n; ++k) cur_mp[k] = (!matched[k] || matched[k] == blocker) ? -1 : index[matched[k]];
    }
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, matches) (src/ORBmatcher.cc:160-289)
int refm_search_by_bow(const Kp* kf_kps, const uint8_t* kf_desc, int n_kf, const uint8_t* kf_has_mp, const int32_t* kf_node_ids,
                       const int32_t* kf_start, const int32_t* kf_items, int kf_nn, const Kp* f_kps, const uint8_t* f_desc, int n_f,
                       const int32_t* f_node_ids, const int32_t* f_start, const int32_t* f_items, int f_nn, float nn_ratio, int check_ori,
                       int32_t* f_mp) {
    FrameView kv{kf_kps, kf_desc, n_kf, 0.f, 0.f, 64.f / 1280.f, 48.f / 720.f, FRAME_GRID_COLS, FRAME_GRID_ROWS};
    FrameView fv{f_kps, f_desc, n_f, 0.f, 0.f, 64.f / 1280.f, 48.f / 720.f, FRAME_GRID_COLS, FRAME_GRID_ROWS};
    set_front_statics(kv);
    Frame K, F;
    fill_front(K, kv); fill_front(F, fv);
    Pool pool;
    std::map<MapPoint*, int> index;
    std::vector<uint8_t> zero(32, 0);
    for (int i = 0; i < n_kf; ++i)
        if (kf_has_mp[i]) { MapPoint* p = make_mp(0, 0, zero.data(), 0, 1); pool.mps.push_back(p); K.mvpMapPoints[i] = p; index[p] = i; }
    for (int a = 0; a < kf_nn; ++a)
        for (int p = kf_start[a]; p < kf_start[a + 1]; ++p) K.mFeatVec[kf_node_ids[a]].push_back((unsigned)kf_items[p]);
    for (int b = 0; b < f_nn; ++b)
        for (int p = f_start[b]; p < f_start[b + 1]; ++p) F.mFeatVec[f_node_ids[b]].push_back((unsigned)f_items[p]);
    KeyFrame* kf = new KeyFrame(K, NULL, NULL);
    pool.kfs.push_back(kf);
    std::vector<MapPoint*> out;
    ORBmatcher m(nn_ratio, check_ori != 0);
    const int n = m.SearchByBoW(kf, F, out);
    for (int k = 0; k < n_f; ++k) f_mp[k] = out[k] ? index[out[k]] : -1;
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (src/ORBmatcher.cc:523-656)
int refm_search_by_bow_kf(const Kp* k1, const uint8_t* d1, int n1, const uint8_t* has1, const int32_t* ids1, const int32_t* st1,
                          const int32_t* it1, int nn1, const Kp* k2, const uint8_t* d2, int n2, const uint8_t* has2, const int32_t* ids2,
                          const int32_t* st2, const int32_t* it2, int nn2, float nn_ratio, int check_ori, int32_t* matches12) {
    FrameView v1{k1, d1, n1, 0.f, 0.f, 64.f / 1280.f, 48.f / 720.f, FRAME_GRID_COLS, FRAME_GRID_ROWS};
    FrameView v2{k2, d2, n2, 0.f, 0.f, 64.f / 1280.f, 48.f / 720.f, FRAME_GRID_COLS, FRAME_GRID_ROWS};
    set_front_statics(v1);
    Frame F1, F2;
    fill_front(F1, v1); fill_front(F2, v2);
    Pool pool;
    std::map<MapPoint*, int> index2;
    std::vector<uint8_t> zero(32, 0);
    for (int i = 0; i < n1; ++i) if (has1[i]) { MapPoint* p = make_mp(0, 0, zero.data(), 0, 1); pool.mps.push_back(p); F1.mvpMapPoints[i] = p; }
    for (int i = 0; i < n2; ++i) if (has2[i]) { MapPoint* p = make_mp(0, 0, zero.data(), 0, 1); pool.mps.push_back(p); F2.mvpMapPoints[i] = p; index2[p] = i; }
    for (int a = 0; a < nn1; ++a) for (int p = st1[a]; p < st1[a + 1]; ++p) F1.mFeatVec[ids1[a]].push_back((unsigned)it1[p]);
    for (int b = 0; b < nn2; ++b) for (int p = st2[b]; p < st2[b + 1]; ++p) F2.mFeatVec[ids2[b]].push_back((unsigned)it2[p]);
    KeyFrame* kf1 = new KeyFrame(F1, NULL, NULL);
    KeyFrame* kf2 = new KeyFrame(F2, NULL, NULL);
    pool.kfs.push_back(kf1); pool.kfs.push_back(kf2);
    std::vector<MapPoint*> out;
    ORBmatcher m(nn_ratio, check_ori != 0);
    const int n = m.SearchByBoW(kf1, kf2, out);
    for (int i = 0; i < n1; ++i) matches12[i] = out[i] ? index2[out[i]] : -1;
    return n;
}

// ORBmatcher::Fuse, both overloads (src/ORBmatcher.cc:826-976 with th, :978-1101 with Scw = identity).  Map points sit at
// (u, v, 1) in front of an identity camera; occupants of key-frame slots are separate map points.  Replace is a logging
// no-op here, so the map state only evolves through AddObservation / AddMapPoint -- the restated oracle models the same.
// act: 0 nothing, 1 pMP->Replace(pMPinKF), 2 pMPinKF->Replace(pMP), 3 added to the empty slot, 5 vpReplacePoint (overload 2).
int refm_fuse(const FrameView* kfv, const float* uright, const float* scale_factors, const float* inv_sigma2, const float* proj,
              const int32_t* level, const uint8_t* mp_desc, const int32_t* mp_nobs, const uint8_t* mp_bad, const uint8_t* mp_in_kf, int n,
              const int32_t* occ, const int32_t* occ_nobs, const uint8_t* occ_bad, int n_occ, float th, int overload, int32_t* act,
              int32_t* slot) {
    if (!check_grid(*kfv, FRAME_GRID_COLS, FRAME_GRID_ROWS)) return -1;
    set_front_statics(*kfv);
    Pool pool;
    Frame K;
    fill_front(K, *kfv, scale_factors);
    K.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + kLevels);
    for (int k = 0; k < kfv->n; ++k) K.mvuRight[k] = uright ? uright[k] : -1.f;
    std::vector<uint8_t> zero(32, 0);
    std::vector<MapPoint*> occmp(n_occ);
    for (int j = 0; j < n_occ; ++j) { occmp[j] = make_mp(0, 0, zero.data(), 0, occ_nobs[j]); occmp[j]->mbBad = occ_bad[j] != 0; pool.mps.push_back(occmp[j]); }
    for (int k = 0; k < kfv->n; ++k) K.mvpMapPoints[k] = occ[k] >= 0 ? occmp[occ[k]] : static_cast<MapPoint*>(NULL);
    KeyFrame* kf = new KeyFrame(K, NULL, NULL);
    pool.kfs.push_back(kf);
    std::vector<MapPoint*> pts(n, static_cast<MapPoint*>(NULL));
    std::map<MapPoint*, int> index;
    for (int i = 0; i < n; ++i) {
        act[i] = 0; slot[i] = -1;
        if (std::isnan(proj[2 * i])) continue;          // overload 1 tolerates NULL entries; overload 2 gets a compacted list
        MapPoint* p = make_mp(proj[2 * i], proj[2 * i + 1], mp_desc + (size_t)i * 32, level[i], mp_nobs[i]);
        p->mbBad = mp_bad[i] != 0;
        if (mp_in_kf[i]) p->mObservations[kf] = 0;
        pool.mps.push_back(p); pts[i] = p; index[p] = i;
    }
    g_replace_log.clear();
    ORBmatcher m(0.6f, true);
    int nFused;
    std::vector<MapPoint*> list, repl;
    std::vector<int> list_i;
    if (overload == 1) {
        nFused = m.Fuse(kf, pts, th);
    } else {
        for (int i = 0; i < n; ++i) if (pts[i]) { list.push_back(pts[i]); list_i.push_back(i); }
        repl.assign(list.size(), static_cast<MapPoint*>(NULL));
        nFused = m.Fuse(kf, cv::Mat::eye(4, 4, CV_32F), list, th, repl);
    }
    std::map<MapPoint*, int> slot_of;
    for (int k = 0; k < kfv->n; ++k) if (kf->mvpMapPoints[k]) slot_of[kf->mvpMapPoints[k]] = k;
    for (int k = 0; k < kfv->n; ++k) {
        MapPoint* p = kf->mvpMapPoints[k];
        if (p && index.count(p)) { act[index[p]] = 3; slot[index[p]] = k; }
    }
    for (size_t e = 0; e < g_replace_log.size(); ++e) {
        MapPoint *a = g_replace_log[e].first, *b = g_replace_log[e].second;
        const bool a_occ = slot_of.count(a) != 0 && !(index.count(a) && act[index[a]] != 3);
        // exactly one of the two sits in a key-frame slot: the occupant; the other is the point being fused
        if (slot_of.count(b) && index.count(a) && !(slot_of.count(a))) { act[index[a]] = 1; slot[index[a]] = slot_of[b]; }
        else if (slot_of.count(a) && index.count(b)) { act[index[b]] = 2; slot[index[b]] = slot_of[a]; }
        (void)a_occ;
    }
    for (size_t e = 0; e < repl.size(); ++e)
        if (repl[e]) { act[list_i[e]] = 5; slot[list_i[e]] = slot_of[repl[e]]; }
    return nFused;
}

// ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1103-1327).  Both key frames sit at the world origin (identity poses), the
// relative transform is s12 = 1, R12 = I, t12 = (tx, ty, 0): a key-frame-1 point at (u, v, 1) projects to (u - tx, v - ty)
// in key frame 2 and vice versa.  pre12[i1] >= 0 pre-fills vpMatches12 (already matched, :1131-1141).
int refm_search_by_sim3(const FrameView* v1, const FrameView* v2, const float* scale_factors, const float* pos1, const int32_t* lvl1,
                        const uint8_t* desc1, const float* pos2, const int32_t* lvl2, const uint8_t* desc2, const int32_t* pre12, float tx,
                        float ty, float th, int32_t* matches12) {
    set_front_statics(*v1);
    Pool pool;
    Frame F1, F2;
    fill_front(F1, *v1, scale_factors); fill_front(F2, *v2, scale_factors);
    std::map<MapPoint*, int> index2;
    for (int i = 0; i < v1->n; ++i)
        if (!std::isnan(pos1[2 * i])) { MapPoint* p = make_mp(pos1[2 * i], pos1[2 * i + 1], desc1 + (size_t)i * 32, lvl1[i], 1); pool.mps.push_back(p); F1.mvpMapPoints[i] = p; }
    for (int i = 0; i < v2->n; ++i)
        if (!std::isnan(pos2[2 * i])) { MapPoint* p = make_mp(pos2[2 * i], pos2[2 * i + 1], desc2 + (size_t)i * 32, lvl2[i], 1); pool.mps.push_back(p); F2.mvpMapPoints[i] = p; index2[p] = i; }
    KeyFrame* kf1 = new KeyFrame(F1, NULL, NULL);
    KeyFrame* kf2 = new KeyFrame(F2, NULL, NULL);
    pool.kfs.push_back(kf1); pool.kfs.push_back(kf2);
    kf1->SetPose(cv::Mat::eye(4, 4, CV_32F)); kf2->SetPose(cv::Mat::eye(4, 4, CV_32F));
    std::vector<MapPoint*> m12(v1->n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < v1->n; ++i)
        if (pre12[i] >= 0 && kf2->mvpMapPoints[pre12[i]]) { m12[i] = kf2->mvpMapPoints[pre12[i]]; m12[i]->mObservations[kf2] = (size_t)pre12[i]; }
    cv::Mat R12 = cv::Mat::eye(3, 3, CV_32F), t12(3, 1, CV_32F);
    t12.at<float>(0) = tx; t12.at<float>(1) = ty; t12.at<float>(2) = 0.f;
    ORBmatcher m(0.75f, true);
    const float s12 = 1.0f;
    const int n = m.SearchBySim3(kf1, kf2, m12, s12, R12, t12, th);
    for (int i = 0; i < v1->n; ++i) matches12[i] = m12[i] ? index2[m12[i]] : -1;
    return n;
}

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:658-824).  The camera is the identity and GetCameraCenter() of
// key frame 1 is the origin, so the epipole of :666-672 is t2w.xy / t2w.z: key frame 2 gets t2w = (ex, ey, 1).
int refm_search_for_triangulation(const Kp* k1, const uint8_t* d1, int n1, const uint8_t* has_mp1, const uint8_t* stereo1,
                                  const int32_t* ids1, const int32_t* st1, const int32_t* it1, int nn1, const Kp* k2, const uint8_t* d2,
                                  int n2, const uint8_t* has_mp2, const uint8_t* stereo2, const int32_t* ids2, const int32_t* st2,
                                  const int32_t* it2, int nn2, const float* F12, float ex, float ey, const float* scale2,
                                  const float* sigma2, int only_stereo, float nn_ratio, int check_ori, int32_t* matches12) {
    FrameView v1{k1, d1, n1, 0.f, 0.f, 64.f / 1280.f, 48.f / 720.f, FRAME_GRID_COLS, FRAME_GRID_ROWS};
    FrameView v2{k2, d2, n2, 0.f, 0.f, 64.f / 1280.f, 48.f / 720.f, FRAME_GRID_COLS, FRAME_GRID_ROWS};
    set_front_statics(v1);
    Frame F1, F2;
    fill_front(F1, v1, scale2); fill_front(F2, v2, scale2);
    F2.mvLevelSigma2.assign(sigma2, sigma2 + kLevels); F1.mvLevelSigma2 = F2.mvLevelSigma2;
    for (int i = 0; i < n1; ++i) F1.mvuRight[i] = stereo1[i] ? 1.f : -1.f;
    for (int i = 0; i < n2; ++i) F2.mvuRight[i] = stereo2[i] ? 1.f : -1.f;
    Pool pool;
    std::vector<uint8_t> zero(32, 0);
    for (int i = 0; i < n1; ++i) if (has_mp1[i]) { MapPoint* p = make_mp(0, 0, zero.data(), 0, 1); pool.mps.push_back(p); F1.mvpMapPoints[i] = p; }
    for (int i = 0; i < n2; ++i) if (has_mp2[i]) { MapPoint* p = make_mp(0, 0, zero.data(), 0, 1); pool.mps.push_back(p); F2.mvpMapPoints[i] = p; }
    for (int a = 0; a < nn1; ++a) for (int p = st1[a]; p < st1[a + 1]; ++p) F1.mFeatVec[ids1[a]].push_back((unsigned)it1[p]);
    for (int b = 0; b < nn2; ++b) for (int p = st2[b]; p < st2[b + 1]; ++p) F2.mFeatVec[ids2[b]].push_back((unsigned)it2[p]);
    KeyFrame* kf1 = new KeyFrame(F1, NULL, NULL);
    KeyFrame* kf2 = new KeyFrame(F2, NULL, NULL);
    pool.kfs.push_back(kf1); pool.kfs.push_back(kf2);
    cv::Mat T2 = cv::Mat::eye(4, 4, CV_32F);
    T2.at<float>(0, 3) = ex; T2.at<float>(1, 3) = ey; T2.at<float>(2, 3) = 1.f;
    kf2->SetPose(T2);
    cv::Mat F(3, 3, CV_32F);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F.at<float>(r, c) = F12[3 * r + c];
    std::vector<std::pair<size_t, size_t> > pairs;
    ORBmatcher m(nn_ratio, check_ori != 0);
    const int n = m.SearchForTriangulation(kf1, kf2, F, pairs, only_stereo != 0);
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    for (size_t k = 0; k < pairs.size(); ++k) matches12[pairs[k].first] = (int)pairs[k].second;
    return n;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-307).  Each observed descriptor sits in a key frame of
// its own; the key frames are constructed in one array so that std::map<KeyFrame*, size_t> walks them in input order.
void refm_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int npts, int32_t* best) {
    Kp kp = {10.f, 10.f, 31.f, 0.f, 1.f, 0, -1};
    for (int p = 0; p < npts; ++p) {
        const int N = start[p + 1] - start[p];
        best[p] = -1;
        if (N == 0) continue;
        const uint8_t* d0 = desc + (size_t)start[p] * 32;
        KeyFrame* kfs = static_cast<KeyFrame*>(::operator new(sizeof(KeyFrame) * (size_t)N));
        std::vector<uint8_t> two(64);
        MapPoint* mp = make_mp(0, 0, d0, 0, 0);
        for (int i = 0; i < N; ++i) {
            // two features per key frame: a decoy in row 0 and the observation in row 1 (exercises `mDescriptors.row(idx)`)
            for (int b = 0; b < 32; ++b) { two[b] = (uint8_t)(~d0[(size_t)i * 32 + b]); two[32 + b] = d0[(size_t)i * 32 + b]; }
            Kp k2[2] = {kp, kp};
            FrameView v{k2, two.data(), 2, 0.f, 0.f, 64.f / 1280.f, 48.f / 720.f, FRAME_GRID_COLS, FRAME_GRID_ROWS};
            set_front_statics(v);
            Frame F;
            fill_front(F, v);
            new (&kfs[i]) KeyFrame(F, NULL, NULL);
            kfs[i].mbBad = false;
            mp->mObservations[&kfs[i]] = 1;
        }
        mp->ComputeDistinctiveDescriptors();
        for (int i = 0; i < N && best[p] < 0; ++i)      // first input row equal to the chosen descriptor
            if (std::memcmp(mp->mDescriptor.ptr(0), d0 + (size_t)i * 32, 32) == 0) best[p] = i;
        for (int i = 0; i < N; ++i) kfs[i].~KeyFrame();
        ::operator delete(kfs);
        delete mp;
    }
}

// Initializer::CheckHomography / CheckFundamental (src/Initializer.cc:391-554) for K hypotheses, one call each
void refm_check_models(const Kp* k1, int n1, const Kp* k2, int n2, const int32_t* matches, int n, const float* A, const float* B, int K,
                       float sigma, int homography, float* scores, uint8_t* inliers) {
    Frame dummy;
    Initializer init(dummy, sigma, 200);
    init.mvKeys1 = to_kps(k1, n1); init.mvKeys2 = to_kps(k2, n2);
    for (int i = 0; i < n; ++i) init.mvMatches12.push_back(std::make_pair(matches[2 * i], matches[2 * i + 1]));
    for (int k = 0; k < K; ++k) {
        cv::Mat M(3, 3, CV_32F), Minv(3, 3, CV_32F);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { M.at<float>(r, c) = A[k * 9 + 3 * r + c]; if (B) Minv.at<float>(r, c) = B[k * 9 + 3 * r + c]; }
        std::vector<bool> in;
        scores[k] = homography ? init.CheckHomography(M, Minv, in, sigma) : init.CheckFundamental(M, in, sigma);
        for (int i = 0; i < n; ++i) inliers[(size_t)k * n + i] = in[i];
    }
}

// Frame::GuidenceKeyBirdPts (src/Frame.cc:671-684: genEdgesPC + the nearEdges filter) on a real Frame object: the verbatim
// build runs the reference's own bodies, the drop-in build runs host/Frame_fbe.cc (filter on the GPU, genEdgesPC unchanged).
// out receives mvKeysBird; counts = {mvKeysBird.size(), mEdgeFree.size(), mEdgeSign.size()}.
void refm_guidance_key_bird_pts(const uint8_t* contour, int rows, int cols, const Kp* kps, int n, Kp* out, int* counts) {
    Frame F;
    F.mBirdviewContourICP = cv::Mat(rows, cols, CV_8U);
    for (int y = 0; y < rows; ++y) std::memcpy(F.mBirdviewContourICP.ptr(y), contour + (size_t)y * cols, (size_t)cols);
    std::vector<cv::KeyPoint> pre = to_kps(kps, n);
    F.GuidenceKeyBirdPts(pre);
    counts[0] = (int)F.mvKeysBird.size(); counts[1] = (int)F.mEdgeFree.size(); counts[2] = (int)F.mEdgeSign.size();
    if (counts[0]) std::memcpy(out, F.mvKeysBird.data(), (size_t)counts[0] * sizeof(Kp));
}

#ifdef FBE_DROPIN
// Frame::UndistortKeyPoints through the drop-in body (host/Frame_fbe.cc); only the drop-in build has it -- the reference's
// own body calls cv::fisheye::undistortPoints, which no shim here provides (its arithmetic is pinned by tests/test_undistort.py).
void refm_undistort_keypoints(const Kp* kps, int n, const float* K, const float* D, Kp* out) {
    Frame F;
    F.N = n;
    F.mvKeys = to_kps(kps, n);
    F.mK = cv::Mat::eye(3, 3, CV_32F);
    F.mK.at<float>(0, 0) = K[0]; F.mK.at<float>(1, 1) = K[1]; F.mK.at<float>(0, 2) = K[2]; F.mK.at<float>(1, 2) = K[3];
    F.mDistCoef = cv::Mat(4, 1, CV_32F);
    for (int i = 0; i < 4; ++i) F.mDistCoef.at<float>(i) = D[i];
    F.UndistortKeyPoints();
    if (n) std::memcpy(out, F.mvKeysUn.data(), (size_t)n * sizeof(Kp));
}

// The bird feature block of the Frame constructor (src/Frame.cc:336-355) through the drop-in helper of host/Frame_fbe.cc, on
// cv::Mat inputs like the constructor's members.  Returns the number of bird keypoints (at most cap are written).
int refm_bird_features(const uint8_t* img, const uint8_t* mask, const uint8_t* contour, int rows, int cols, Kp* out, uint8_t* desc, int cap) {
    cv::Mat I(rows, cols, CV_8U), M, Cn;
    for (int y = 0; y < rows; ++y) std::memcpy(I.ptr(y), img + (size_t)y * cols, (size_t)cols);
    if (mask) { M = cv::Mat(rows, cols, CV_8U); for (int y = 0; y < rows; ++y) std::memcpy(M.ptr(y), mask + (size_t)y * cols, (size_t)cols); }
    if (contour) { Cn = cv::Mat(rows, cols, CV_8U); for (int y = 0; y < rows; ++y) std::memcpy(Cn.ptr(y), contour + (size_t)y * cols, (size_t)cols); }
    std::vector<cv::KeyPoint> keys;
    cv::Mat D;
    FbeBirdFeatures(I, M, Cn, keys, D);
    const int n = (int)keys.size();
    for (int i = 0; i < n && i < cap; ++i) { std::memcpy(out + i, &keys[i], sizeof(Kp)); std::memcpy(desc + (size_t)i * 32, D.ptr(i), 32); }
    return n;
}
#endif

int refm_hamming256(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b)); }

}  // extern "C"
