// Drop-in bodies for the ORBmatcher methods on the hot path.  This file is compiled INSIDE the reference tree in place
// of the same-named method bodies of src/ORBmatcher.cc (see INTEGRATION.md for the exact patch); the class declaration
// include/ORBmatcher.h is unchanged, so Tracking.cc / LocalMapping.cc / LoopClosing.cc call it as before.  Methods that
// are not searches (ComputeThreeMaxima, RadiusByViewingCos, CheckDistEpipolarLine) keep the reference's own CPU code.
//
// In the build image of this repository it is compiled against the reference's own headers with OpenCV replaced by the
// test shim oracle/cvshim_m (oracle/Makefile target `dropinmatch`) and checked on the GPU against the outputs of the
// reference's verbatim CPU build (tests/test_gpu_dropin_match.py).
//
// Everything pointer-valued stays on the host: frames are flattened to (keypoints, descriptors, grid geometry),
// map points to (projection, level, descriptor) arrays; the C-ABI returns index lists which are turned back into
// pointer writes here.  cv::Mat arithmetic of the reference (projections, bird pixel conversion, the distance
// filter of BirdMapPointMatch) is executed here exactly as the reference writes it.
#include <climits>
#include <cmath>
#include <limits>
#include <set>
#include <vector>

#include "Converter.h"
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "MapPointBird.h"
#include "ORBmatcher.h"
#include "fbe_cabi.h"
#include "fbe_host.h"

namespace ORB_SLAM2 {

namespace {

static_assert(sizeof(cv::KeyPoint) == sizeof(fbe_keypoint), "cv::KeyPoint must be 28 bytes");

struct Matcher {          // one device matcher per calling thread and per (ratio, orientation) setting
    fbe_matcher* h;
    float ratio;
    bool ori;
};

// A C-ABI failure (no GPU, CUDA error, capacity, invalid arguments) is fatal, exactly as in the extractor and Frame shims:
// returning "0 matches" would read as "tracking lost" instead of a fault.  There is deliberately no CPU fallback.
void die(const char* what, int rc) {
    fprintf(stderr, "ORBmatcher (fbe-b200): %s failed: %d (%s)\n", what, rc, fbe_last_error());
    abort();
}
#define FBE_CK(call) do { const int _rc = (call); if (_rc != FBE_OK) die(#call, _rc); } while (0)

struct MatcherCache {     // the handles of a thread are destroyed when the thread exits
    std::vector<Matcher> v;
    ~MatcherCache() { for (size_t i = 0; i < v.size(); ++i) fbe_matcher_destroy(v[i].h); }
};

fbe_matcher* matcher_for(float ratio, bool ori) {
    static thread_local MatcherCache cache;
    for (size_t i = 0; i < cache.v.size(); ++i)
        if (cache.v[i].ratio == ratio && cache.v[i].ori == ori) return cache.v[i].h;
    Matcher m = {NULL, ratio, ori};
    FBE_CK(fbe_matcher_create(ratio, ori ? 1 : 0, fbe_host_device(), &m.h));     // same GPU as the extractor (FBE_DEVICE)
    cache.v.push_back(m);
    return m.h;
}

const unsigned char* desc_ptr(const cv::Mat& d) { return d.empty() ? NULL : d.ptr<unsigned char>(0); }

fbe_frame_view front_view(const Frame& F) {
    fbe_frame_view v;
    v.kps = reinterpret_cast<const fbe_keypoint*>(F.mvKeysUn.data());
    v.desc = desc_ptr(F.mDescriptors);         // rows are contiguous: created by OutputArray::create(n, 32, CV_8U)
    v.n = F.N;
    v.min_x = Frame::mnMinX; v.min_y = Frame::mnMinY;
    v.inv_w = Frame::mfGridElementWidthInv; v.inv_h = Frame::mfGridElementHeightInv;
    v.gcols = FRAME_GRID_COLS; v.grows = FRAME_GRID_ROWS;
    return v;
}

fbe_frame_view bird_view(const Frame& F) {
    fbe_frame_view v;
    v.kps = reinterpret_cast<const fbe_keypoint*>(F.mvKeysBird.data());
    v.desc = desc_ptr(F.mDescriptorsBird);
    v.n = (int)F.mvKeysBird.size();
    v.min_x = 0.f; v.min_y = 0.f;
    v.inv_w = Frame::mfGridElementWidthInvBirdview; v.inv_h = Frame::mfGridElementHeightInvBirdview;
    v.gcols = FRAME_GRID_BIRD; v.grows = FRAME_GRID_BIRD;
    return v;
}

void copy_desc(const cv::Mat& d, unsigned char* dst) { std::memcpy(dst, d.ptr<unsigned char>(0), 32); }

}  // namespace

// src/ORBmatcher.cc:406-521
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                        std::vector<int>& vnMatches12, int windowSize) {
    vnMatches12 = std::vector<int>(F1.mvKeysUn.size(), -1);
    fbe_frame_view v1 = front_view(F1), v2 = front_view(F2);
    int nmatches = 0;
    static_assert(sizeof(cv::Point2f) == 8, "Point2f layout");
    FBE_CK(fbe_search_for_initialization(matcher_for(mfNNratio, mbCheckOrientation), &v1, &v2,
                                  reinterpret_cast<float*>(vbPrevMatched.data()), vnMatches12.data(), windowSize, &nmatches));
    return nmatches;
}

// src/ORBmatcher.cc:1329-1471 (monocular: bForward/bBackward are false when bMono)
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const int n = LastFrame.N;
    std::vector<float> proj(2 * (size_t)n, std::numeric_limits<float>::quiet_NaN());
    std::vector<unsigned char> mpdesc(32 * (size_t)n, 0), has_obs(n, 1);
    for (int i = 0; i < n; i++) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw * x3Dw + tcw;
        const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
        const float invzc = 1.0 / x3Dc.at<float>(2);
        if (invzc < 0) continue;
        float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
        float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
        if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
        proj[2 * i] = u; proj[2 * i + 1] = v;
        copy_desc(pMP->GetDescriptor(), &mpdesc[32 * (size_t)i]);
        has_obs[i] = pMP->Observations() > 0;
    }
    std::vector<unsigned char> taken(CurrentFrame.N, 0);
    for (int k = 0; k < CurrentFrame.N; k++)
        if (CurrentFrame.mvpMapPoints[k] && CurrentFrame.mvpMapPoints[k]->Observations() > 0) taken[k] = 1;
    std::vector<int> cur_mp(CurrentFrame.N, -1);
    fbe_frame_view cv_ = front_view(CurrentFrame);
    int nmatches = 0;
    if (!bMono) {   // the stereo branches (bForward / bBackward level windows :1349-1350,1386-1390, mvuRight check :1417-1424) are not
                    // built: this fork tracks monocular + bird view, and silently running the mono search would change results
        fprintf(stderr, "ORBmatcher::SearchByProjection (fbe-b200): bMono == false (stereo / RGB-D) is not supported by the drop-in\n");
        abort();
    }
    FBE_CK(fbe_search_by_projection_last(matcher_for(mfNNratio, mbCheckOrientation), &cv_,
                                  reinterpret_cast<const fbe_keypoint*>(LastFrame.mvKeysUn.data()), proj.data(), mpdesc.data(), n,
                                  CurrentFrame.mvScaleFactors.data(), (int)CurrentFrame.mvScaleFactors.size(), taken.data(),
                                  has_obs.data(), th, cur_mp.data(), &nmatches));
    // pointer writes of :1431 and :1461: -2 marks keypoints assigned by this call and then removed by the orientation
    // histogram, where the reference writes NULL whatever the keypoint held before
    for (int k = 0; k < CurrentFrame.N; k++) {
        if (cur_mp[k] >= 0) CurrentFrame.mvpMapPoints[k] = LastFrame.mvpMapPoints[cur_mp[k]];
        else if (cur_mp[k] == -2) CurrentFrame.mvpMapPoints[k] = static_cast<MapPoint*>(NULL);
    }
    return nmatches;
}

// src/ORBmatcher.cc:1473-1600 (relocalisation)
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                                   const int ORBdist) {
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const int n = (int)vpMPs.size();
    std::vector<float> proj(2 * (size_t)n, std::numeric_limits<float>::quiet_NaN());
    std::vector<int> level(n, 0);
    std::vector<unsigned char> mpdesc(32 * (size_t)n, 0);
    for (int i = 0; i < n; i++) {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw * x3Dw + tcw;
        const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
        const float invzc = 1.0 / x3Dc.at<float>(2);
        const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
        const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
        if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
        cv::Mat PO = x3Dw - Ow;
        const float dist3D = cv::norm(PO);
        if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
        level[i] = pMP->PredictScale(dist3D, &CurrentFrame);
        proj[2 * i] = u; proj[2 * i + 1] = v;
        copy_desc(pMP->GetDescriptor(), &mpdesc[32 * (size_t)i]);
    }
    std::vector<unsigned char> taken(CurrentFrame.N, 0);
    for (int k = 0; k < CurrentFrame.N; k++)
        if (CurrentFrame.mvpMapPoints[k]) taken[k] = 1;                 // :1541
    std::vector<int> cur_mp(CurrentFrame.N, -1);
    fbe_frame_view cv_ = front_view(CurrentFrame);
    int nmatches = 0;
    FBE_CK(fbe_search_by_projection_reloc(matcher_for(mfNNratio, mbCheckOrientation), &cv_,
                                   reinterpret_cast<const fbe_keypoint*>(pKF->mvKeysUn.data()), proj.data(), level.data(),
                                   mpdesc.data(), n, CurrentFrame.mvScaleFactors.data(), (int)CurrentFrame.mvScaleFactors.size(),
                                   taken.data(), th, ORBdist, cur_mp.data(), &nmatches));
    for (int k = 0; k < CurrentFrame.N; k++) {
        if (cur_mp[k] >= 0) CurrentFrame.mvpMapPoints[k] = vpMPs[cur_mp[k]];
        else if (cur_mp[k] == -2) CurrentFrame.mvpMapPoints[k] = static_cast<MapPoint*>(NULL);
    }
    return nmatches;
}

// src/ORBmatcher.cc:291-404 (loop closing)
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints,
                                   std::vector<MapPoint*>& vpMatched, int th) {
    const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy;
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
    const int n = (int)vpPoints.size();
    std::vector<float> proj(2 * (size_t)n, std::numeric_limits<float>::quiet_NaN());
    std::vector<int> level(n, 0);
    std::vector<unsigned char> mpdesc(32 * (size_t)n, 0);
    for (int i = 0; i < n; i++) {
        MapPoint* pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = Rcw * p3Dw + tcw;
        if (p3Dc.at<float>(2) < 0.0) continue;
        const float invz = 1 / p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0) * invz, y = p3Dc.at<float>(1) * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!pKF->IsInImage(u, v)) continue;
        cv::Mat PO = p3Dw - Ow;
        const float dist = cv::norm(PO);
        if (dist < pMP->GetMinDistanceInvariance() || dist > pMP->GetMaxDistanceInvariance()) continue;
        cv::Mat Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist) continue;
        level[i] = pMP->PredictScale(dist, pKF);
        proj[2 * i] = u; proj[2 * i + 1] = v;
        copy_desc(pMP->GetDescriptor(), &mpdesc[32 * (size_t)i]);
    }
    const int N = (int)pKF->mvKeysUn.size();
    std::vector<unsigned char> matched(N, 0);
    for (int k = 0; k < N; k++)
        if (vpMatched[k]) matched[k] = 1;                               // :371
    std::vector<int> kf_mp(N, -1);
    fbe_frame_view kv;
    kv.kps = reinterpret_cast<const fbe_keypoint*>(pKF->mvKeysUn.data());
    kv.desc = desc_ptr(pKF->mDescriptors);
    kv.n = N;
    kv.min_x = pKF->mnMinX; kv.min_y = pKF->mnMinY;
    kv.inv_w = pKF->mfGridElementWidthInv; kv.inv_h = pKF->mfGridElementHeightInv;
    kv.gcols = pKF->mnGridCols; kv.grows = pKF->mnGridRows;
    int nmatches = 0;
    FBE_CK(fbe_search_by_projection_loop(matcher_for(mfNNratio, mbCheckOrientation), &kv, proj.data(), level.data(), mpdesc.data(), n,
                                  pKF->mvScaleFactors.data(), (int)pKF->mvScaleFactors.size(), matched.data(), th, kf_mp.data(),
                                  &nmatches));
    for (int k = 0; k < N; k++)
        if (kf_mp[k] >= 0) vpMatched[k] = vpPoints[kf_mp[k]];
    return nmatches;
}

// src/ORBmatcher.cc:46-130
int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    std::vector<int> src;
    std::vector<float> proj, viewcos;
    std::vector<int> level;
    std::vector<unsigned char> desc, has_obs;
    for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
        MapPoint* pMP = vpMapPoints[iMP];
        if (!pMP->mbTrackInView || pMP->isBad()) continue;
        src.push_back((int)iMP);
        proj.push_back(pMP->mTrackProjX); proj.push_back(pMP->mTrackProjY);
        level.push_back(pMP->mnTrackScaleLevel);
        viewcos.push_back(pMP->mTrackViewCos);
        desc.resize(desc.size() + 32);
        copy_desc(pMP->GetDescriptor(), &desc[desc.size() - 32]);
        has_obs.push_back(pMP->Observations() > 0);
    }
    std::vector<unsigned char> taken(F.N, 0);
    for (int k = 0; k < F.N; k++)
        if (F.mvpMapPoints[k] && F.mvpMapPoints[k]->Observations() > 0) taken[k] = 1;
    std::vector<int> cur_mp(F.N, -1);
    fbe_frame_view v = front_view(F);
    int nmatches = 0;
    FBE_CK(fbe_search_by_projection_map(matcher_for(mfNNratio, mbCheckOrientation), &v, F.mvScaleFactors.data(), (int)F.mvScaleFactors.size(),
                                 proj.data(), level.data(), viewcos.data(), desc.data(), (int)src.size(), taken.data(),
                                 has_obs.data(), th, cur_mp.data(), &nmatches));
    for (int k = 0; k < F.N; k++)
        if (cur_mp[k] >= 0) F.mvpMapPoints[k] = vpMapPoints[src[cur_mp[k]]];
    return nmatches;
}

namespace {
// DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned int>> -> CSR over ascending node ids
struct Csr { std::vector<int> ids, start, items; };
Csr flatten(const DBoW2::FeatureVector& fv) {
    Csr c;
    c.start.push_back(0);
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        c.ids.push_back((int)it->first);
        c.items.insert(c.items.end(), it->second.begin(), it->second.end());
        c.start.push_back((int)c.items.size());
    }
    return c;
}
std::vector<unsigned char> good_points(const std::vector<MapPoint*>& v) {
    std::vector<unsigned char> has(v.size(), 0);
    for (size_t i = 0; i < v.size(); i++) has[i] = v[i] && !v[i]->isBad();
    return has;
}
}  // namespace

// src/ORBmatcher.cc:160-289
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    const std::vector<unsigned char> has_mp = good_points(vpMapPointsKF);
    Csr a = flatten(pKF->mFeatVec), b = flatten(F.mFeatVec);
    std::vector<int> f_mp(F.N, -1);
    int nmatches = 0;
    FBE_CK(fbe_search_by_bow(matcher_for(mfNNratio, mbCheckOrientation), reinterpret_cast<const fbe_keypoint*>(pKF->mvKeysUn.data()),
                      desc_ptr(pKF->mDescriptors), (int)pKF->mvKeysUn.size(), has_mp.data(), a.ids.data(), a.start.data(),
                      a.items.data(), (int)a.ids.size(), reinterpret_cast<const fbe_keypoint*>(F.mvKeys.data()), desc_ptr(F.mDescriptors),
                      F.N, b.ids.data(), b.start.data(), b.items.data(), (int)b.ids.size(), f_mp.data(), &nmatches));
    for (int k = 0; k < F.N; k++)
        if (f_mp[k] >= 0) vpMapPointMatches[k] = vpMapPointsKF[f_mp[k]];
    return nmatches;
}

// src/ORBmatcher.cc:523-656 (loop closing: key frame against key frame)
int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    const std::vector<unsigned char> has1 = good_points(vpMapPoints1), has2 = good_points(vpMapPoints2);
    Csr a = flatten(pKF1->mFeatVec), b = flatten(pKF2->mFeatVec);
    std::vector<int> m12(std::max<size_t>(vpMapPoints1.size(), 1), -1);
    int nmatches = 0;
    FBE_CK(fbe_search_by_bow_kf(matcher_for(mfNNratio, mbCheckOrientation), reinterpret_cast<const fbe_keypoint*>(pKF1->mvKeysUn.data()),
                         desc_ptr(pKF1->mDescriptors), (int)vpMapPoints1.size(), has1.data(), a.ids.data(), a.start.data(), a.items.data(),
                         (int)a.ids.size(), reinterpret_cast<const fbe_keypoint*>(pKF2->mvKeysUn.data()), desc_ptr(pKF2->mDescriptors),
                         (int)vpMapPoints2.size(), has2.data(), b.ids.data(), b.start.data(), b.items.data(), (int)b.ids.size(),
                         m12.data(), &nmatches));
    for (size_t i = 0; i < vpMapPoints1.size(); i++)
        if (m12[i] >= 0) vpMatches12[i] = vpMapPoints2[m12[i]];
    return nmatches;
}

// src/ORBmatcher.cc:658-824 (LocalMapping::CreateNewMapPoints)
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
                                       std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
    // epipole in the second image, exactly as :666-672
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    const float ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
    const float ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;

    const int n1 = pKF1->N, n2 = pKF2->N;
    std::vector<unsigned char> skip1(std::max(n1, 1)), st1(std::max(n1, 1)), skip2(std::max(n2, 1)), st2(std::max(n2, 1));
    for (int i = 0; i < n1; i++) {
        st1[i] = pKF1->mvuRight[i] >= 0;
        skip1[i] = pKF1->GetMapPoint(i) != NULL || (bOnlyStereo && !st1[i]);
    }
    for (int i = 0; i < n2; i++) {
        st2[i] = pKF2->mvuRight[i] >= 0;
        skip2[i] = pKF2->GetMapPoint(i) != NULL || (bOnlyStereo && !st2[i]);
    }
    Csr a = flatten(pKF1->mFeatVec), b = flatten(pKF2->mFeatVec);
    float F[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) F[3 * r + c] = F12.at<float>(r, c);
    std::vector<int> m12(std::max(n1, 1), -1);
    int nmatches = 0;
    FBE_CK(fbe_search_for_triangulation(matcher_for(mfNNratio, mbCheckOrientation), reinterpret_cast<const fbe_keypoint*>(pKF1->mvKeysUn.data()),
                                 desc_ptr(pKF1->mDescriptors), n1, skip1.data(), st1.data(), a.ids.data(), a.start.data(), a.items.data(),
                                 (int)a.ids.size(), reinterpret_cast<const fbe_keypoint*>(pKF2->mvKeysUn.data()),
                                 desc_ptr(pKF2->mDescriptors), n2, skip2.data(), st2.data(), b.ids.data(), b.start.data(), b.items.data(),
                                 (int)b.ids.size(), F, ex, ey, pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(),
                                 (int)pKF2->mvScaleFactors.size(), m12.data(), &nmatches));
    vMatchedPairs.clear();
    vMatchedPairs.reserve(nmatches);
    for (int i = 0; i < n1; i++)
        if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));
    return nmatches;
}

namespace {
fbe_frame_view keyframe_view(KeyFrame* pKF) {
    fbe_frame_view kv;
    kv.kps = reinterpret_cast<const fbe_keypoint*>(pKF->mvKeysUn.data());
    kv.desc = desc_ptr(pKF->mDescriptors);
    kv.n = (int)pKF->mvKeysUn.size();
    kv.min_x = pKF->mnMinX; kv.min_y = pKF->mnMinY;
    kv.inv_w = pKF->mfGridElementWidthInv; kv.inv_h = pKF->mfGridElementHeightInv;
    kv.gcols = pKF->mnGridCols; kv.grows = pKF->mnGridRows;
    return kv;
}
}  // namespace

// src/ORBmatcher.cc:826-976.  The candidate search runs on the device for every point that is eligible on entry; the map
// updates are then replayed in order on the host, re-checking isBad() / IsInKeyFrame() at that moment (an eligible point
// can only become ineligible through the replay, never the other way round; the search does not read map state).
int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    cv::Mat Rcw = pKF->GetRotation();
    cv::Mat tcw = pKF->GetTranslation();
    const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy, &bf = pKF->mbf;
    cv::Mat Ow = pKF->GetCameraCenter();
    const int nMPs = (int)vpMapPoints.size();
    std::vector<float> proj(2 * (size_t)std::max(nMPs, 1), std::numeric_limits<float>::quiet_NaN()), pur(std::max(nMPs, 1), 0.f), radius(std::max(nMPs, 1), 0.f);
    std::vector<int> level(std::max(nMPs, 1), 0);
    std::vector<unsigned char> mpdesc(32 * (size_t)std::max(nMPs, 1), 0);
    for (int i = 0; i < nMPs; i++) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = Rcw * p3Dw + tcw;
        if (p3Dc.at<float>(2) < 0.0f) continue;
        const float invz = 1 / p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0) * invz, y = p3Dc.at<float>(1) * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!pKF->IsInImage(u, v)) continue;
        const float ur = u - bf * invz;
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        cv::Mat PO = p3Dw - Ow;
        const float dist3D = cv::norm(PO);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        cv::Mat Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist3D) continue;
        const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
        level[i] = nPredictedLevel;
        radius[i] = th * pKF->mvScaleFactors[nPredictedLevel];
        proj[2 * i] = u; proj[2 * i + 1] = v; pur[i] = ur;
        copy_desc(pMP->GetDescriptor(), &mpdesc[32 * (size_t)i]);
    }
    std::vector<int> best_idx(std::max(nMPs, 1), -1), best_dist(std::max(nMPs, 1), INT_MAX);
    fbe_frame_view kv = keyframe_view(pKF);
    FBE_CK(fbe_fuse_search(matcher_for(mfNNratio, mbCheckOrientation), &kv, pKF->mvuRight.data(), pKF->mvInvLevelSigma2.data(),
                    (int)pKF->mvInvLevelSigma2.size(), proj.data(), pur.data(), level.data(), radius.data(), mpdesc.data(), nMPs, 1,
                    best_idx.data(), best_dist.data()));
    int nFused = 0;
    for (int i = 0; i < nMPs; i++) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP || proj[2 * i] != proj[2 * i]) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        if (best_dist[i] <= TH_LOW) {
            const int bestIdx = best_idx[i];
            MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) {
                    if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                    else pMPinKF->Replace(pMP);
                }
            } else {
                pMP->AddObservation(pKF, bestIdx);
                pKF->AddMapPoint(pMP, bestIdx);
            }
            nFused++;
        }
    }
    return nFused;
}

// src/ORBmatcher.cc:978-1101 (loop closing)
int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
    const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy;
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
    const int nPoints = (int)vpPoints.size();
    std::vector<float> proj(2 * (size_t)std::max(nPoints, 1), std::numeric_limits<float>::quiet_NaN()), radius(std::max(nPoints, 1), 0.f);
    std::vector<int> level(std::max(nPoints, 1), 0);
    std::vector<unsigned char> mpdesc(32 * (size_t)std::max(nPoints, 1), 0);
    for (int iMP = 0; iMP < nPoints; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = Rcw * p3Dw + tcw;
        if (p3Dc.at<float>(2) < 0.0f) continue;
        const float invz = 1.0 / p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0) * invz, y = p3Dc.at<float>(1) * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!pKF->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        cv::Mat PO = p3Dw - Ow;
        const float dist3D = cv::norm(PO);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        cv::Mat Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist3D) continue;
        const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
        level[iMP] = nPredictedLevel;
        radius[iMP] = th * pKF->mvScaleFactors[nPredictedLevel];
        proj[2 * iMP] = u; proj[2 * iMP + 1] = v;
        copy_desc(pMP->GetDescriptor(), &mpdesc[32 * (size_t)iMP]);
    }
    std::vector<int> best_idx(std::max(nPoints, 1), -1), best_dist(std::max(nPoints, 1), INT_MAX);
    fbe_frame_view kv = keyframe_view(pKF);
    FBE_CK(fbe_fuse_search(matcher_for(mfNNratio, mbCheckOrientation), &kv, NULL, NULL, 0, proj.data(), NULL, level.data(), radius.data(),
                    mpdesc.data(), nPoints, 0, best_idx.data(), best_dist.data()));
    int nFused = 0;
    for (int iMP = 0; iMP < nPoints; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (proj[2 * iMP] != proj[2 * iMP]) continue;
        if (pMP->isBad()) continue;                      // spAlreadyFound is a snapshot taken on entry (:998), already applied
        if (best_dist[iMP] <= TH_LOW) {
            const int bestIdx = best_idx[iMP];
            MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF;
            } else {
                pMP->AddObservation(pKF, bestIdx);
                pKF->AddMapPoint(pMP, bestIdx);
            }
            nFused++;
        }
    }
    return nFused;
}

namespace {
// one direction of SearchBySim3 (:1152-1222 / :1225-1295): points of `src` projected into `dst` with dst_from_src(p3Dc_src)
void sim3_direction(ORBmatcher* self, fbe_matcher* m, KeyFrame* src, KeyFrame* dst, const std::vector<MapPoint*>& pts,
                    const std::vector<bool>& already, const cv::Mat& Rsw, const cv::Mat& tsw, const cv::Mat& sRds, const cv::Mat& tds,
                    float th, std::vector<int>& match) {
    (void)self; (void)src;
    const float &fx = dst->fx, &fy = dst->fy, &cx = dst->cx, &cy = dst->cy;     // same camera for both key frames
    const int N = (int)pts.size();
    std::vector<float> proj(2 * (size_t)std::max(N, 1), std::numeric_limits<float>::quiet_NaN()), radius(std::max(N, 1), 0.f);
    std::vector<int> level(std::max(N, 1), 0);
    std::vector<unsigned char> mpdesc(32 * (size_t)std::max(N, 1), 0);
    for (int i = 0; i < N; i++) {
        MapPoint* pMP = pts[i];
        if (!pMP || already[i]) continue;
        if (pMP->isBad()) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dcs = Rsw * p3Dw + tsw;
        cv::Mat p3Dcd = sRds * p3Dcs + tds;
        if (p3Dcd.at<float>(2) < 0.0) continue;
        const float invz = 1.0 / p3Dcd.at<float>(2);
        const float x = p3Dcd.at<float>(0) * invz, y = p3Dcd.at<float>(1) * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!dst->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        const float dist3D = cv::norm(p3Dcd);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = pMP->PredictScale(dist3D, dst);
        level[i] = nPredictedLevel;
        radius[i] = th * dst->mvScaleFactors[nPredictedLevel];
        proj[2 * i] = u; proj[2 * i + 1] = v;
        copy_desc(pMP->GetDescriptor(), &mpdesc[32 * (size_t)i]);
    }
    std::vector<int> best_idx(std::max(N, 1), -1), best_dist(std::max(N, 1), INT_MAX);
    fbe_frame_view kv = keyframe_view(dst);
    FBE_CK(fbe_fuse_search(m, &kv, NULL, NULL, 0, proj.data(), NULL, level.data(), radius.data(), mpdesc.data(), N, 0, best_idx.data(),
                    best_dist.data()));
    match.assign(N, -1);
    for (int i = 0; i < N; i++)
        if (best_dist[i] <= ORBmatcher::TH_HIGH) match[i] = best_idx[i];
}
}  // namespace

// src/ORBmatcher.cc:1103-1327 (loop closing): the two directional searches use the Fuse candidate search (same level band,
// no reprojection gate, strict best distance); the agreement check stays on the host.
int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th) {
    cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation();
    cv::Mat R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12 * R12;
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {
        MapPoint* pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    fbe_matcher* m = matcher_for(mfNNratio, mbCheckOrientation);
    std::vector<int> vnMatch1, vnMatch2;
    sim3_direction(this, m, pKF1, pKF2, vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, th, vnMatch1);
    sim3_direction(this, m, pKF2, pKF1, vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, th, vnMatch2);
    int nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
    }
    return nFound;
}

// src/ORBmatcher.cc:1602-1760, isProject == 0 (every call site of the reference passes 0)
int ORBmatcher::BirdviewMatch(Frame& CurF, const std::vector<cv::KeyPoint>& vRefKeysBird, const cv::Mat& DescriptorsBird,
                              const std::vector<MapPointBird*>& vRefMapPointsBird, std::vector<cv::DMatch>& vDMatches12,
                              int isProject, int windowSize) {
    (void)vRefMapPointsBird;
    if (isProject) { fprintf(stderr, "BirdviewMatch(isProject=1) is not on the accelerated path\n"); abort(); }
    const int n = (int)vRefKeysBird.size();
    std::vector<int> dm(3 * (size_t)std::max(n, 1));
    int nd = 0, nmatches = 0;
    fbe_frame_view v = bird_view(CurF);
    FBE_CK(fbe_birdview_match(matcher_for(mfNNratio, mbCheckOrientation), reinterpret_cast<const fbe_keypoint*>(vRefKeysBird.data()),
                       desc_ptr(DescriptorsBird), n, &v, windowSize, dm.data(), &nd, &nmatches));
    for (int i = 0; i < nd; i++) vDMatches12.push_back(cv::DMatch(dm[3 * i], dm[3 * i + 1], (float)dm[3 * i + 2]));
    return nmatches;
}

// src/ORBmatcher.cc:1763-1902
int ORBmatcher::BirdMapPointMatch(Frame& CurF, const std::vector<MapPointBird*>& vRefMapPointsBird, int windowSize, float filterSize) {
    const cv::Mat Tbw = Frame::Tbc * CurF.mTcw;
    const int n = (int)vRefMapPointsBird.size();
    std::vector<float> pix(2 * (size_t)n, std::numeric_limits<float>::quiet_NaN());
    std::vector<unsigned char> desc(32 * (size_t)n, 0);
    for (int i1 = 0; i1 < n; i1++) {
        MapPointBird* p = vRefMapPointsBird[i1];
        if (!p) continue;
        cv::Mat worldPos = p->GetWorldPos();
        cv::Mat localPos = Tbw.rowRange(0, 3).colRange(0, 3) * worldPos + Tbw.rowRange(0, 3).col(3);
        if (fabs(localPos.at<float>(2)) > 0.2) continue;
        cv::Point2f pt = Converter::BaseXY2BirdPixel(cv::Point3f(localPos.at<float>(0), localPos.at<float>(1), localPos.at<float>(2)));
        if (pt.x < 0 || pt.x >= Frame::birdviewCols || pt.y < 0 || pt.y >= Frame::birdviewRows) continue;
        pix[2 * i1] = pt.x; pix[2 * i1 + 1] = pt.y;
        copy_desc(p->GetDescriptor(), &desc[32 * (size_t)i1]);
    }
    std::vector<int> vnMatches12(n, -1);
    int nmatches = 0;
    fbe_frame_view v = bird_view(CurF);
    FBE_CK(fbe_bird_map_point_match(matcher_for(mfNNratio, mbCheckOrientation), pix.data(), desc.data(), n, &v, windowSize,
                             vnMatches12.data(), &nmatches));
    // second pass, :1865-1895, verbatim semantics (host arithmetic, `> 0` quirk, last writer wins)
    int InlierMatches = 0;
    cv::Mat Tcw2 = CurF.mTcw;
    for (int i1 = 0; i1 < n; i1++) {
        if (vnMatches12[i1] > 0) {
            MapPointBird* p = vRefMapPointsBird[i1];
            if (!p) continue;
            cv::Mat ptwC = p->GetWorldPos();
            cv::Mat ptc2c = Tcw2.rowRange(0, 3).colRange(0, 3) * ptwC + Tcw2.rowRange(0, 3).col(3);
            cv::Mat pt2c(CurF.mvKeysBirdCamXYZ[vnMatches12[i1]]);
            double disC = cv::norm(ptc2c - pt2c, cv::NORM_L2);
            if (disC < filterSize) {
                CurF.mvpMapPointsBird[vnMatches12[i1]] = p;
                InlierMatches++;
            }
        }
    }
    return InlierMatches;
}

// src/ORBmatcher.cc:1951-1967 -- stays a host inline (called pairwise from MapPoint.cc:281, MapPointBird.cc:129, Frame.cc:872)
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    return fbe_hamming256(a.ptr<unsigned char>(), b.ptr<unsigned char>());
}

}  // namespace ORB_SLAM2
