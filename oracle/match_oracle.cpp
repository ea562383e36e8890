// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement, on plain arrays, of the reference's keypoint grid and of the named ORBmatcher searches.
// These functions follow the reference line by line on POD arrays (file:line below, relative to /root/reference).
// PINNED: the reference ships no tests or golden vectors for these paths, so the pin is the reference's own
// compiled code -- src/ORBmatcher.cc built VERBATIM against a header-only cv::Mat shim, together with the verbatim
// Frame::AssignFeaturesToGrid / GetFeaturesInArea[Birdview], KeyFrame::GetFeaturesInArea, MapPoint::PredictScale and
// Converter::BaseXY2BirdPixel (oracle/Makefile target `refmatch`, oracle/gen_ref_parts.py, oracle/ref_match_wrap.cpp ->
// oracle/_ref/libfbe_refmatch.so).  tests/test_oracle_vs_refmatch.py demands equality with that build on seeded scenes
// and, on machines without it, with its committed outputs (tests/golden/match.npz); tests/test_oracle_match.py adds an
// independent pure-Python restatement and hand-built cases for each documented quirk.
//   Frame::AssignFeaturesToGrid / PosInGrid / PosInGridBirdview   src/Frame.cc:381-411, 548-570
//   Frame::GetFeaturesInArea / GetFeaturesInAreaBirdview           src/Frame.cc:493-546, 572-626
//   ORBmatcher::DescriptorDistance                                 src/ORBmatcher.cc:1951-1967
//   ORBmatcher::ComputeThreeMaxima                                 src/ORBmatcher.cc:1905-1946
//   ORBmatcher::SearchForInitialization                            src/ORBmatcher.cc:406-521
//   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) src/ORBmatcher.cc:46-130
//   ORBmatcher::SearchByProjection(Frame&, const Frame&, th, mono) src/ORBmatcher.cc:1329-1471
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...)                src/ORBmatcher.cc:160-289
//   ORBmatcher::BirdviewMatch (isProject == 0)                     src/ORBmatcher.cc:1602-1760
//   ORBmatcher::BirdMapPointMatch (first pass)                     src/ORBmatcher.cc:1763-1863
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;

struct Kp { float x, y, size, angle, response; int32_t octave, class_id; };

struct FrameView {
    const Kp* kps;
    const uint8_t* desc;
    int32_t n;
    float min_x, min_y, inv_w, inv_h;
    int32_t gcols, grows;
};

struct Grid {
    int gcols, grows;
    std::vector<std::vector<int> > cell;   // [ix*grows + iy], push_back order
};

Grid build_grid(const FrameView& f) {
    Grid g;
    g.gcols = f.gcols; g.grows = f.grows;
    g.cell.assign((size_t)f.gcols * f.grows, std::vector<int>());
    for (int i = 0; i < f.n; ++i) {
        int px = (int)std::round((f.kps[i].x - f.min_x) * f.inv_w);
        int py = (int)std::round((f.kps[i].y - f.min_y) * f.inv_h);
        if (px < 0 || px >= f.gcols || py < 0 || py >= f.grows) continue;
        g.cell[(size_t)px * f.grows + py].push_back(i);
    }
    return g;
}

// upper_inclusive = true : Frame::GetFeaturesInArea        (ix <= nMaxCellX)
// upper_inclusive = false: Frame::GetFeaturesInAreaBirdview (ix <  nMaxCellX)   -- quirk Q2
std::vector<int> features_in_area(const FrameView& f, const Grid& g, float x, float y, float r, int minLevel,
                                  int maxLevel, bool upper_inclusive) {
    std::vector<int> out;
    const int nMinCellX = std::max(0, (int)std::floor((x - f.min_x - r) * f.inv_w));
    if (nMinCellX >= g.gcols) return out;
    const int nMaxCellX = std::min(g.gcols - 1, (int)std::ceil((x - f.min_x + r) * f.inv_w));
    if (nMaxCellX < 0) return out;
    const int nMinCellY = std::max(0, (int)std::floor((y - f.min_y - r) * f.inv_h));
    if (nMinCellY >= g.grows) return out;
    const int nMaxCellY = std::min(g.grows - 1, (int)std::ceil((y - f.min_y + r) * f.inv_h));
    if (nMaxCellY < 0) return out;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    const int ex = upper_inclusive ? 1 : 0;
    for (int ix = nMinCellX; ix < nMaxCellX + ex; ++ix)
        for (int iy = nMinCellY; iy < nMaxCellY + ex; ++iy) {
            const std::vector<int>& c = g.cell[(size_t)ix * g.grows + iy];
            for (size_t j = 0; j < c.size(); ++j) {
                const Kp& kp = f.kps[c[j]];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float dx = kp.x - x, dy = kp.y - y;
                if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(c[j]);
            }
        }
    return out;
}

int hamming(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t pa, pb;
        std::memcpy(&pa, a + 4 * i, 4);
        std::memcpy(&pb, b + 4 * i, 4);
        unsigned v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

int rot_bin(float a1, float a2) {
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

}  // namespace

extern "C" {

int orc_hamming256(const uint8_t* a, const uint8_t* b) { return hamming(a, b); }

int orc_grid_assign(const Kp* kps, int n, float min_x, float min_y, float inv_w, float inv_h, int gcols, int grows,
                    int32_t* cell_start, int32_t* cell_items) {
    FrameView f{kps, nullptr, n, min_x, min_y, inv_w, inv_h, gcols, grows};
    Grid g = build_grid(f);
    int off = 0;
    for (int c = 0; c < gcols * grows; ++c) {
        cell_start[c] = off;
        for (size_t j = 0; j < g.cell[c].size(); ++j) cell_items[off++] = g.cell[c][j];
    }
    cell_start[gcols * grows] = off;
    return off;
}

// candidate list of one query, for direct tests of the area queries
int orc_features_in_area(const FrameView* f, float x, float y, float r, int minLevel, int maxLevel, int upper_inclusive,
                         int32_t* out, int cap) {
    Grid g = build_grid(*f);
    std::vector<int> v = features_in_area(*f, g, x, y, r, minLevel, maxLevel, upper_inclusive != 0);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}

int orc_search_for_initialization(const FrameView* F1, const FrameView* F2, float* prev_matched, int32_t* matches12,
                                  int window, float nn_ratio, int check_ori) {
    int nmatches = 0;
    for (int i = 0; i < F1->n; ++i) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> vMatchedDistance(F2->n, INT_MAX), vnMatches21(F2->n, -1);
    Grid g2 = build_grid(*F2);
    for (int i1 = 0; i1 < F1->n; ++i1) {
        const Kp& kp1 = F1->kps[i1];
        const int level1 = kp1.octave;
        if (level1 > 0) continue;
        std::vector<int> cand = features_in_area(*F2, g2, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)window,
                                                 level1, level1, true);
        if (cand.empty()) continue;
        const uint8_t* d1 = F1->desc + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (size_t k = 0; k < cand.size(); ++k) {
            const int i2 = cand[k];
            const int dist = hamming(d1, F2->desc + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nn_ratio) {
                if (vnMatches21[bestIdx2] >= 0) { matches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (check_ori) rotHist[rot_bin(F1->kps[i1].angle, F2->kps[bestIdx2].angle)].push_back(i1);
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); ++j) {
                const int idx1 = rotHist[i][j];
                if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
            }
        }
    }
    for (int i1 = 0; i1 < F1->n; ++i1)
        if (matches12[i1] >= 0) {
            prev_matched[2 * i1] = F2->kps[matches12[i1]].x;
            prev_matched[2 * i1 + 1] = F2->kps[matches12[i1]].y;
        }
    return nmatches;
}

// BirdviewMatch, isProject == 0.  dmatches: up to n_ref triples (queryIdx, trainIdx, distance).
int orc_birdview_match(const Kp* ref_kps, const uint8_t* ref_desc, int n_ref, const FrameView* cur, int window,
                       float nn_ratio, int check_ori, int32_t* dmatches, int32_t* n_dmatches) {
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> vnMatches21(cur->n, -1), vnMatches12(n_ref, -1), vMatchedDistance(n_ref, INT_MAX);
    Grid g = build_grid(*cur);
    for (int i1 = 0; i1 < n_ref; ++i1) {
        const Kp& kp1 = ref_kps[i1];
        const int level1 = kp1.octave;
        if (level1 > 0) continue;
        std::vector<int> cand = features_in_area(*cur, g, kp1.x, kp1.y, (float)window, level1, level1, false);
        if (cand.empty()) continue;
        const uint8_t* d1 = ref_desc + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx = -1;
        for (size_t k = 0; k < cand.size(); ++k) {
            const int i2 = cand[k];
            if (i2 >= cur->n) continue;
            const int dist = hamming(d1, cur->desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nn_ratio) {
                vnMatches21[bestIdx] = i1;
                vnMatches12[i1] = bestIdx;
                vMatchedDistance[i1] = bestDist;
                nmatches++;
            }
            if (check_ori) rotHist[rot_bin(kp1.angle, cur->kps[bestIdx].angle)].push_back(i1);   // even if the ratio test failed (Q7)
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); ++j) {
                const int idx1 = rotHist[i][j];
                if (vnMatches12[idx1] >= 0) { vnMatches21[vnMatches12[idx1]] = -1; vnMatches12[idx1] = -1; nmatches--; }
            }
        }
    }
    int nd = 0;
    for (int i = 0; i < n_ref; ++i)
        if (vnMatches12[i] > 0) {   // index 0 is dropped (Q8)
            dmatches[3 * nd] = i; dmatches[3 * nd + 1] = vnMatches12[i]; dmatches[3 * nd + 2] = vMatchedDistance[i];
            ++nd;
        }
    *n_dmatches = nd;
    return nmatches;
}

// BirdMapPointMatch, first pass (:1763-1863).  mp_pix: n x 2 projected bird pixels, NaN x = map point skipped
// (NULL pointer, |z| > 0.2 or outside the image -- decided by the caller with the reference's host arithmetic).
int orc_bird_map_point_match(const float* mp_pix, const uint8_t* mp_desc, int n_mp, const FrameView* cur, int window,
                             float nn_ratio, int32_t* matches12) {
    int nmatches = 0;
    Grid g = build_grid(*cur);
    for (int i1 = 0; i1 < n_mp; ++i1) {
        matches12[i1] = -1;
        const float px = mp_pix[2 * i1], py = mp_pix[2 * i1 + 1];
        if (std::isnan(px)) continue;
        std::vector<int> cand = features_in_area(*cur, g, px, py, (float)window, -1, -1, false);
        if (cand.empty()) continue;
        const uint8_t* d1 = mp_desc + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx = -1;
        for (size_t k = 0; k < cand.size(); ++k) {
            const int i2 = cand[k];
            if (i2 >= cur->n) continue;
            const int dist = hamming(d1, cur->desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW && bestDist < (float)bestDist2 * nn_ratio) { matches12[i1] = bestIdx; nmatches++; }
    }
    return nmatches;
}

// SearchByProjection(CurrentFrame, LastFrame, th, bMono = true).  last_proj: n_last x 2 projected (u,v), NaN u =
// skipped (no map point, outlier, behind the camera or outside the image bounds -- caller's host arithmetic).
// cur_taken: 1 where CurrentFrame.mvpMapPoints[k] has Observations() > 0 on entry.  last_has_obs: 1 where the
// last-frame map point has Observations() > 0 (so that, once assigned, it blocks its keypoint); NULL = all.
int orc_search_by_projection_last(const FrameView* cur, const Kp* last_kps, const float* last_proj,
                                  const uint8_t* last_mp_desc, int n_last, const float* scale_factors,
                                  const uint8_t* cur_taken, const uint8_t* last_has_obs, float th, int check_ori,
                                  int32_t* cur_mp) {
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    Grid g = build_grid(*cur);
    std::vector<uint8_t> taken(cur->n, 0);
    for (int k = 0; k < cur->n; ++k) { cur_mp[k] = -1; taken[k] = cur_taken ? cur_taken[k] : 0; }
    for (int i = 0; i < n_last; ++i) {
        const float u = last_proj[2 * i], v = last_proj[2 * i + 1];
        if (std::isnan(u)) continue;
        const int nLastOctave = last_kps[i].octave;
        const float radius = th * scale_factors[nLastOctave];
        std::vector<int> cand = features_in_area(*cur, g, u, v, radius, nLastOctave - 1, nLastOctave + 1, true);
        if (cand.empty()) continue;
        const uint8_t* dMP = last_mp_desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (size_t k = 0; k < cand.size(); ++k) {
            const int i2 = cand[k];
            if (taken[i2]) continue;
            const int dist = hamming(dMP, cur->desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            cur_mp[bestIdx2] = i;
            if (!last_has_obs || last_has_obs[i]) taken[bestIdx2] = 1;
            nmatches++;
            if (check_ori) rotHist[rot_bin(last_kps[i].angle, cur->kps[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); ++j) { cur_mp[rotHist[i][j]] = -2; nmatches--; }   // -2: the reference writes NULL here
    }
    return nmatches;
}

// SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)  src/ORBmatcher.cc:1473-1600 (relocalisation) and
// SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)               src/ORBmatcher.cc:291-404   (loop closing).
// Both project map points into a frame / key frame, search a radius th * scale[predicted level], take the single best
// unmatched candidate and accept it below a distance bound; they differ in the level band (+1 / +0 above the
// predicted level), the bound (ORBdist / TH_LOW) and the orientation histogram (reloc only).  The projection, the
// bounds / depth / viewing-angle rejections and MapPoint::PredictScale (src/MapPoint.cc:385-417) stay with the caller:
// proj NaN = rejected, level = nPredictedLevel.  q_kps supplies the key-frame keypoint angle of each map point (reloc).
// The loop variant filters levels inside the candidate loop after KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:901-940,
// same cell arithmetic as the Frame version, no level filter); filtering during the walk visits the same candidates in
// the same order.
int orc_search_by_projection_kf(const FrameView* cur, const Kp* q_kps, const float* proj, const int32_t* level,
                                const uint8_t* mp_desc, int n_mp, const float* scale_factors, const uint8_t* cur_taken,
                                float th, int th_dist, int level_up, int check_ori, int32_t* cur_mp) {
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    Grid g = build_grid(*cur);
    std::vector<uint8_t> taken(cur->n, 0);
    for (int k = 0; k < cur->n; ++k) { cur_mp[k] = -1; taken[k] = cur_taken ? cur_taken[k] : 0; }
    for (int i = 0; i < n_mp; ++i) {
        const float u = proj[2 * i], v = proj[2 * i + 1];
        if (std::isnan(u)) continue;
        const int nPredictedLevel = level[i];
        const float radius = th * scale_factors[nPredictedLevel];
        std::vector<int> cand = features_in_area(*cur, g, u, v, radius, nPredictedLevel - 1, nPredictedLevel + level_up, true);
        if (cand.empty()) continue;
        const uint8_t* dMP = mp_desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (size_t k = 0; k < cand.size(); ++k) {
            const int i2 = cand[k];
            if (taken[i2]) continue;
            const int dist = hamming(dMP, cur->desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= th_dist && bestIdx2 >= 0) {
            cur_mp[bestIdx2] = i;
            taken[bestIdx2] = 1;
            nmatches++;
            if (check_ori) rotHist[rot_bin(q_kps[i].angle, cur->kps[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); ++j) { cur_mp[rotHist[i][j]] = -2; nmatches--; }   // -2: the reference writes NULL here
    }
    return nmatches;
}

// SearchByProjection(F, vpMapPoints, th).  Map points already filtered by mbTrackInView && !isBad().
int orc_search_by_projection_map(const FrameView* cur, const float* scale_factors, const float* mp_proj,
                                 const int32_t* mp_level, const float* mp_viewcos, const uint8_t* mp_desc, int n_mp,
                                 const uint8_t* cur_taken, const uint8_t* mp_has_obs, float th, float nn_ratio,
                                 int32_t* cur_mp) {
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    Grid g = build_grid(*cur);
    std::vector<uint8_t> taken(cur->n, 0);
    for (int k = 0; k < cur->n; ++k) { cur_mp[k] = -1; taken[k] = cur_taken ? cur_taken[k] : 0; }
    for (int iMP = 0; iMP < n_mp; ++iMP) {
        const int nPredictedLevel = mp_level[iMP];
        float r = mp_viewcos[iMP] > 0.998 ? 2.5f : 4.0f;
        if (bFactor) r *= th;
        std::vector<int> cand = features_in_area(*cur, g, mp_proj[2 * iMP], mp_proj[2 * iMP + 1],
                                                 r * scale_factors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, true);
        if (cand.empty()) continue;
        const uint8_t* d = mp_desc + (size_t)iMP * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (size_t k = 0; k < cand.size(); ++k) {
            const int idx = cand[k];
            if (taken[idx]) continue;
            const int dist = hamming(d, cur->desc + (size_t)idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = cur->kps[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = cur->kps[idx].octave; bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nn_ratio * bestDist2) continue;
            cur_mp[bestIdx] = iMP;
            if (!mp_has_obs || mp_has_obs[iMP]) taken[bestIdx] = 1;
            nmatches++;
        }
    }
    return nmatches;
}

// SearchByBoW(KeyFrame*, Frame&, matches).  Feature vectors as CSR over ascending node ids.
int orc_search_by_bow(const Kp* kf_kps, const uint8_t* kf_desc, int n_kf, const uint8_t* kf_has_mp,
                      const int32_t* kf_node_ids, const int32_t* kf_start, const int32_t* kf_items, int kf_nn,
                      const Kp* f_kps, const uint8_t* f_desc, int n_f, const int32_t* f_node_ids,
                      const int32_t* f_start, const int32_t* f_items, int f_nn, float nn_ratio, int check_ori,
                      int32_t* f_mp) {
    (void)n_kf;
    for (int k = 0; k < n_f; ++k) f_mp[k] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    int a = 0, b = 0;
    while (a < kf_nn && b < f_nn) {
        if (kf_node_ids[a] == f_node_ids[b]) {
            for (int p = kf_start[a]; p < kf_start[a + 1]; ++p) {
                const int realIdxKF = kf_items[p];
                if (!kf_has_mp[realIdxKF]) continue;
                const uint8_t* dKF = kf_desc + (size_t)realIdxKF * 32;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int q = f_start[b]; q < f_start[b + 1]; ++q) {
                    const int realIdxF = f_items[q];
                    if (f_mp[realIdxF] >= 0) continue;
                    const int dist = hamming(dKF, f_desc + (size_t)realIdxF * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW) {
                    if ((float)bestDist1 < nn_ratio * (float)bestDist2) {
                        f_mp[bestIdxF] = realIdxKF;
                        if (check_ori) rotHist[rot_bin(kf_kps[realIdxKF].angle, f_kps[bestIdxF].angle)].push_back(bestIdxF);
                        nmatches++;
                    }
                }
            }
            ++a; ++b;
        } else if (kf_node_ids[a] < f_node_ids[b]) {
            while (a < kf_nn && kf_node_ids[a] < f_node_ids[b]) ++a;      // lower_bound
        } else {
            while (b < f_nn && f_node_ids[b] < kf_node_ids[a]) ++b;
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); ++j) { f_mp[rotHist[i][j]] = -2; nmatches--; }   // -2: the reference writes NULL here
        }
    }
    return nmatches;
}

// brute-force top-2 (stress config C5): ties -> lowest target index, second counts duplicates
// SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12)  src/ORBmatcher.cc:523-656 (loop closing): both sides key frames,
// `bestDist1 < TH_LOW` strict, vbMatched2 blocks matched key-frame-2 features, result indexed by key-frame-1 feature.
int orc_search_by_bow_kf(const Kp* k1, const uint8_t* d1, int n1, const uint8_t* has1, const int32_t* ids1, const int32_t* st1,
                         const int32_t* it1, int nn1, const Kp* k2, const uint8_t* d2, int n2, const uint8_t* has2,
                         const int32_t* ids2, const int32_t* st2, const int32_t* it2, int nn2, float nn_ratio, int check_ori,
                         int32_t* matches12) {
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<uint8_t> vbMatched2(std::max(n2, 1), 0);
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0, a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (ids1[a] == ids2[b]) {
            for (int p1 = st1[a]; p1 < st1[a + 1]; ++p1) {
                const int idx1 = it1[p1];
                if (!has1[idx1]) continue;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int p2 = st2[b]; p2 < st2[b + 1]; ++p2) {
                    const int idx2 = it2[p2];
                    if (vbMatched2[idx2] || !has2[idx2]) continue;
                    const int dist = hamming(d1 + (size_t)idx1 * 32, d2 + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW && (float)bestDist1 < nn_ratio * (float)bestDist2) {
                    matches12[idx1] = bestIdx2;
                    vbMatched2[bestIdx2] = 1;
                    if (check_ori) rotHist[rot_bin(k1[idx1].angle, k2[bestIdx2].angle)].push_back(idx1);
                    nmatches++;
                }
            }
            ++a; ++b;
        } else if (ids1[a] < ids2[b]) ++a;
        else ++b;
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); ++j) { matches12[rotHist[i][j]] = -1; nmatches--; }
    }
    return nmatches;
}

// Frame::isInFrustum  src/Frame.cc:435-491 with MapPoint::PredictScale(dist, Frame*)  src/MapPoint.cc:402-417 and the
// invariance distances  src/MapPoint.cc:373-383.  The cv::Mat expressions are evaluated the way OpenCV evaluates them
// (pinned against cv2 4.13 by tests/test_frustum.py): `mRcw*P+mtcw` is ONE gemm through the small-matrix path -- float
// products and sums left to right, then (float)((double)t*alpha + (double)c*beta); cv::norm and Mat::dot accumulate in
// double.  `level_boundary` flags points whose predicted level sits within 4 ulp of an integer quotient (the only place
// where a different-but-valid logf may change the answer).
struct FrustumView { float Rcw[9], tcw[3], Ow[3], fx, fy, cx, cy, min_x, max_x, min_y, max_y, mbf, log_scale_factor; int32_t n_levels; };
void orc_is_in_frustum(const FrustumView* v, const float* pos, const float* normal, const float* min_dist, const float* max_dist,
                       int n, float cos_limit, uint8_t* in_view, float* proj, float* proj_xr, int32_t* level, float* view_cos,
                       uint8_t* level_boundary) {
    for (int i = 0; i < n; ++i) {
        in_view[i] = 0; proj[2 * i] = proj[2 * i + 1] = 0.f; proj_xr[i] = 0.f; level[i] = 0; view_cos[i] = 0.f; level_boundary[i] = 0;
        const float* P = pos + 3 * i;
        float Pc[3];
        for (int r = 0; r < 3; ++r) {
            const float t = v->Rcw[3 * r] * P[0] + v->Rcw[3 * r + 1] * P[1] + v->Rcw[3 * r + 2] * P[2];
            Pc[r] = (float)((double)t * 1.0 + (double)v->tcw[r] * 1.0);
        }
        const float PcX = Pc[0], PcY = Pc[1], PcZ = Pc[2];
        if (PcZ < 0.0f) continue;
        const float invz = 1.0f / PcZ;
        const float u = v->fx * PcX * invz + v->cx;
        const float w = v->fy * PcY * invz + v->cy;
        if (u < v->min_x || u > v->max_x) continue;
        if (w < v->min_y || w > v->max_y) continue;
        const float maxDistance = 1.2f * max_dist[i], minDistance = 0.8f * min_dist[i];
        const float PO[3] = {P[0] - v->Ow[0], P[1] - v->Ow[1], P[2] - v->Ow[2]};
        double ss = 0;
        for (int k = 0; k < 3; ++k) ss += (double)PO[k] * (double)PO[k];
        const float dist = (float)std::sqrt(ss);
        if (dist < minDistance || dist > maxDistance) continue;
        double dot = 0;
        for (int k = 0; k < 3; ++k) dot += (double)PO[k] * (double)normal[3 * i + k];
        const float viewCos = (float)(dot / dist);
        if (viewCos < cos_limit) continue;
        const float ratio = max_dist[i] / dist;
        const float q = std::log(ratio) / v->log_scale_factor;                 // float overloads (using namespace std in the reference TU)
        int nScale = (int)std::ceil(q);
        if (nScale < 0) nScale = 0;
        else if (nScale >= v->n_levels) nScale = v->n_levels - 1;
        level_boundary[i] = std::fabs(q - std::nearbyint(q)) <= 4.f * 1.1920929e-7f * std::fmax(1.f, std::fabs(q));
        in_view[i] = 1; proj[2 * i] = u; proj[2 * i + 1] = w; proj_xr[i] = u - v->mbf * invz; level[i] = nScale; view_cos[i] = viewCos;
    }
}

// The candidate search shared by both ORBmatcher::Fuse overloads (src/ORBmatcher.cc:893-948, :1053-1079): walk
// KeyFrame::GetFeaturesInArea(u, v, radius) (src/KeyFrame.cc:901-940, no level filter, inclusive cell bounds), keep levels
// [nPredictedLevel-1, nPredictedLevel], optionally gate on the reprojection error (:911-937), strict `dist<bestDist`.
static void fuse_search_one(const FrameView* kf, const Grid& g, const float* uright, const float* inv_sigma2, float u, float v, float ur,
                            int nPredictedLevel, float radius, const uint8_t* dMP, int check_chi2, int& bestIdx, int& bestDist) {
    bestIdx = -1; bestDist = INT_MAX;
    std::vector<int> vIndices = features_in_area(*kf, g, u, v, radius, -1, -1, true);
    for (size_t c = 0; c < vIndices.size(); ++c) {
        const int idx = vIndices[c];
        const Kp& kp = kf->kps[idx];
        const int kpLevel = kp.octave;
        if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
        if (check_chi2) {
            if (uright && uright[idx] >= 0) {
                const float ex = u - kp.x, ey = v - kp.y, er = ur - uright[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * inv_sigma2[kpLevel] > 7.8) continue;
            } else {
                const float ex = u - kp.x, ey = v - kp.y;
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_sigma2[kpLevel] > 5.99) continue;
            }
        }
        const int dist = hamming(dMP, kf->desc + (size_t)idx * 32);
        if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
}

void orc_fuse_search(const FrameView* kf, const float* uright, const float* inv_sigma2, const float* proj, const float* proj_ur,
                     const int32_t* level, const float* radius, const uint8_t* mp_desc, int n, int check_chi2, int32_t* best_idx,
                     int32_t* best_dist) {
    Grid g = build_grid(*kf);
    for (int i = 0; i < n; ++i) {
        best_idx[i] = -1; best_dist[i] = INT_MAX;
        if (std::isnan(proj[2 * i])) continue;
        fuse_search_one(kf, g, uright, inv_sigma2, proj[2 * i], proj[2 * i + 1], proj_ur ? proj_ur[i] : 0.f, level[i], radius[i],
                        mp_desc + (size_t)i * 32, check_chi2, best_idx[i], best_dist[i]);
    }
}

// ORBmatcher::Fuse end to end, overload 1 (:826-976) or 2 (:978-1101), on an abstract map state: per point Observations() /
// isBad() / IsInKeyFrame(pKF), per key-frame slot an occupant id with its own Observations() / isBad().  MapPoint::Replace
// is treated as a recorded no-op (what the verbatim harness does), AddObservation / AddMapPoint update the state.
// act: 0 nothing, 1 pMP->Replace(pMPinKF), 2 pMPinKF->Replace(pMP), 3 added to the empty slot, 5 vpReplacePoint[i] set.
int orc_fuse(const FrameView* kf, const float* uright, const float* inv_sigma2, const float* proj, const float* proj_ur,
             const int32_t* level, const float* radius, const uint8_t* mp_desc, const int32_t* mp_nobs, const uint8_t* mp_bad,
             const uint8_t* mp_in_kf, int n, const int32_t* occ, const int32_t* occ_nobs, const uint8_t* occ_bad, int overload,
             int32_t* act, int32_t* slot) {
    Grid g = build_grid(*kf);
    // slot state: -1 empty, j >= 0 original occupant j, -(i+2) map point i added by this call
    std::vector<int> cur(occ, occ + kf->n);
    std::vector<int> nobs(mp_nobs, mp_nobs + n);
    std::vector<uint8_t> inkf(mp_in_kf, mp_in_kf + n);
    int nFused = 0;
    for (int i = 0; i < n; ++i) {
        act[i] = 0; slot[i] = -1;
        if (std::isnan(proj[2 * i])) continue;
        if (mp_bad[i] || (overload == 1 && inkf[i])) continue;
        int bestIdx, bestDist;
        fuse_search_one(kf, g, uright, inv_sigma2, proj[2 * i], proj[2 * i + 1], proj_ur ? proj_ur[i] : 0.f, level[i], radius[i],
                        mp_desc + (size_t)i * 32, overload == 1, bestIdx, bestDist);
        if (bestDist <= TH_LOW) {
            const int o = cur[bestIdx];
            if (o != -1) {
                const bool o_bad = o >= 0 ? occ_bad[o] != 0 : mp_bad[-o - 2] != 0;
                const int o_nobs = o >= 0 ? occ_nobs[o] : nobs[-o - 2];
                if (!o_bad) {
                    if (overload == 1) act[i] = o_nobs > nobs[i] ? 1 : 2;
                    else act[i] = 5;
                    slot[i] = bestIdx;
                }
            } else {
                nobs[i]++; inkf[i] = 1;                  // AddObservation
                cur[bestIdx] = -(i + 2);                 // AddMapPoint
                act[i] = 3; slot[i] = bestIdx;
            }
            nFused++;
        }
    }
    return nFused;
}

// ORBmatcher::SearchBySim3  src/ORBmatcher.cc:1103-1327 after the projections: proj12 / level12 = key-frame-1 map points in
// key frame 2 (NaN u = no point, already matched, bad, or rejected by depth / image / distance tests), proj21 / level21 the
// other direction; radius = th * mvScaleFactors[level]; pre12 = vpMatches12 on entry as key-frame-2 indices (kept).
int orc_search_by_sim3(const FrameView* kf1, const FrameView* kf2, const float* scale_factors, const float* proj12, const int32_t* level12,
                       const uint8_t* desc1, const float* proj21, const int32_t* level21, const uint8_t* desc2, const int32_t* pre12,
                       float th, int32_t* matches12) {
    Grid g1 = build_grid(*kf1), g2 = build_grid(*kf2);
    std::vector<int> vnMatch1(kf1->n, -1), vnMatch2(kf2->n, -1);
    for (int i1 = 0; i1 < kf1->n; ++i1) {
        if (std::isnan(proj12[2 * i1])) continue;
        int bestIdx, bestDist;
        fuse_search_one(kf2, g2, nullptr, nullptr, proj12[2 * i1], proj12[2 * i1 + 1], 0.f, level12[i1], th * scale_factors[level12[i1]],
                        desc1 + (size_t)i1 * 32, 0, bestIdx, bestDist);
        if (bestDist <= TH_HIGH) vnMatch1[i1] = bestIdx;
    }
    for (int i2 = 0; i2 < kf2->n; ++i2) {
        if (std::isnan(proj21[2 * i2])) continue;
        int bestIdx, bestDist;
        fuse_search_one(kf1, g1, nullptr, nullptr, proj21[2 * i2], proj21[2 * i2 + 1], 0.f, level21[i2], th * scale_factors[level21[i2]],
                        desc2 + (size_t)i2 * 32, 0, bestIdx, bestDist);
        if (bestDist <= TH_HIGH) vnMatch2[i2] = bestIdx;
    }
    int nFound = 0;
    for (int i1 = 0; i1 < kf1->n; ++i1) {
        matches12[i1] = pre12[i1];
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1) { matches12[i1] = idx2; nFound++; }
    }
    return nFound;
}

// ORBmatcher::SearchForTriangulation  src/ORBmatcher.cc:658-824 with CheckDistEpipolarLine :141-158.  skipN = feature has a
// map point or fails the bOnlyStereo filter; stereoN = mvuRight >= 0; (ex, ey) = epipole in key frame 2 (:666-672, computed
// by the caller).  vbMatched2 exists in the reference but is never set, which is reproduced by not having it.
int orc_search_for_triangulation(const Kp* k1, const uint8_t* d1, int n1, const uint8_t* skip1, const uint8_t* stereo1,
                                 const int32_t* ids1, const int32_t* st1, const int32_t* it1, int nn1, const Kp* k2, const uint8_t* d2,
                                 int n2, const uint8_t* skip2, const uint8_t* stereo2, const int32_t* ids2, const int32_t* st2,
                                 const int32_t* it2, int nn2, const float* F12, float ex, float ey, const float* scale2,
                                 const float* sigma2, int check_ori, int32_t* matches12) {
    (void)n2;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0, a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (ids1[a] < ids2[b]) { ++a; continue; }
        if (ids1[a] > ids2[b]) { ++b; continue; }
        for (int p1 = st1[a]; p1 < st1[a + 1]; ++p1) {
            const int idx1 = it1[p1];
            if (skip1[idx1]) continue;
            const bool bStereo1 = stereo1[idx1] != 0;
            const Kp& kp1 = k1[idx1];
            int bestDist = TH_LOW, bestIdx2 = -1;
            for (int p2 = st2[b]; p2 < st2[b + 1]; ++p2) {
                const int idx2 = it2[p2];
                if (skip2[idx2]) continue;
                const int dist = hamming(d1 + (size_t)idx1 * 32, d2 + (size_t)idx2 * 32);
                if (dist > TH_LOW || dist > bestDist) continue;
                const Kp& kp2 = k2[idx2];
                if (!bStereo1 && !stereo2[idx2]) {
                    const float distex = ex - kp2.x, distey = ey - kp2.y;
                    if (distex * distex + distey * distey < 100 * scale2[kp2.octave]) continue;
                }
                // CheckDistEpipolarLine
                const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
                const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
                const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
                const float num = la * kp2.x + lb * kp2.y + lc;
                const float den = la * la + lb * lb;
                if (den == 0) continue;
                const float dsqr = num * num / den;
                if (dsqr < 3.84 * sigma2[kp2.octave]) { bestIdx2 = idx2; bestDist = dist; }
            }
            if (bestIdx2 >= 0) {
                matches12[idx1] = bestIdx2;
                nmatches++;
                if (check_ori) rotHist[rot_bin(kp1.angle, k2[bestIdx2].angle)].push_back(idx1);
            }
        }
        ++a; ++b;
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); ++j) { matches12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// MapPoint::ComputeDistinctiveDescriptors  src/MapPoint.cc:242-307 (MapPointBird.cc:90-155 repeats it): for each map point
// the observed descriptors start[p] .. start[p+1]-1 in the reference's walking order -> index of the chosen one (-1: none).
void orc_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int npts, int32_t* best, int32_t* best_median) {
    for (int p = 0; p < npts; ++p) {
        const size_t N = (size_t)(start[p + 1] - start[p]);
        best[p] = -1; best_median[p] = 0;
        if (N == 0) continue;
        const uint8_t* d0 = desc + (size_t)start[p] * 32;
        std::vector<float> Distances(N * N);
        for (size_t i = 0; i < N; i++) {
            Distances[i * N + i] = 0;
            for (size_t j = i + 1; j < N; j++) {
                const int distij = hamming(d0 + i * 32, d0 + j * 32);
                Distances[i * N + j] = distij;
                Distances[j * N + i] = distij;
            }
        }
        int BestMedian = INT_MAX, BestIdx = 0;
        for (size_t i = 0; i < N; i++) {
            std::vector<int> vDists(Distances.begin() + i * N, Distances.begin() + (i + 1) * N);
            std::sort(vDists.begin(), vDists.end());
            const int median = vDists[0.5 * (N - 1)];
            if (median < BestMedian) { BestMedian = median; BestIdx = (int)i; }
        }
        best[p] = BestIdx; best_median[p] = BestMedian;
    }
}

// Initializer::CheckHomography (src/Initializer.cc:391-474) and CheckFundamental (:476-554) for K hypotheses.  A = H21 or
// F21, B = H12 (homography), row-major 3x3 each.  Compiled with -ffp-contract=off: every operation rounds on its own.
void orc_check_models(const Kp* k1, const Kp* k2, const int32_t* matches, int n, const float* A, const float* B, int K, float sigma,
                      int homography, float* scores, uint8_t* inliers) {
    const float invSigmaSquare = 1.0 / (sigma * sigma);
    for (int k = 0; k < K; ++k) {
        const float* a = A + k * 9;
        const float* b = homography ? B + k * 9 : nullptr;
        float score = 0;
        const float th = homography ? 5.991 : 3.841, thScore = 5.991;
        for (int i = 0; i < n; ++i) {
            const float u1 = k1[matches[2 * i]].x, v1 = k1[matches[2 * i]].y, u2 = k2[matches[2 * i + 1]].x, v2 = k2[matches[2 * i + 1]].y;
            float chiSquare1, chiSquare2;
            if (homography) {
                const float w2in1inv = 1.0 / (b[6] * u2 + b[7] * v2 + b[8]);
                const float u2in1 = (b[0] * u2 + b[1] * v2 + b[2]) * w2in1inv;
                const float v2in1 = (b[3] * u2 + b[4] * v2 + b[5]) * w2in1inv;
                chiSquare1 = ((u1 - u2in1) * (u1 - u2in1) + (v1 - v2in1) * (v1 - v2in1)) * invSigmaSquare;
                const float w1in2inv = 1.0 / (a[6] * u1 + a[7] * v1 + a[8]);
                const float u1in2 = (a[0] * u1 + a[1] * v1 + a[2]) * w1in2inv;
                const float v1in2 = (a[3] * u1 + a[4] * v1 + a[5]) * w1in2inv;
                chiSquare2 = ((u2 - u1in2) * (u2 - u1in2) + (v2 - v1in2) * (v2 - v1in2)) * invSigmaSquare;
            } else {
                const float a2 = a[0] * u1 + a[1] * v1 + a[2], b2 = a[3] * u1 + a[4] * v1 + a[5], c2 = a[6] * u1 + a[7] * v1 + a[8];
                const float num2 = a2 * u2 + b2 * v2 + c2;
                chiSquare1 = num2 * num2 / (a2 * a2 + b2 * b2) * invSigmaSquare;
                const float a1 = a[0] * u2 + a[3] * v2 + a[6], b1 = a[1] * u2 + a[4] * v2 + a[7], c1 = a[2] * u2 + a[5] * v2 + a[8];
                const float num1 = a1 * u1 + b1 * v1 + c1;
                chiSquare2 = num1 * num1 / (a1 * a1 + b1 * b1) * invSigmaSquare;
            }
            bool bIn = true;
            if (chiSquare1 > th) bIn = false; else score += thScore - chiSquare1;
            if (chiSquare2 > th) bIn = false; else score += thScore - chiSquare2;
            inliers[(size_t)k * n + i] = bIn;
        }
        scores[k] = score;
    }
}

// DBoW2 TemplatedVocabulary::transform(feature, word_id, weight, &nid, levelsup)  Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1263
// on a tree given as the rows of the text format (loadFromTextFile :1338-1436): nodes 1..n in creation order, parent id,
// nIsLeaf flag, descriptor, weight.  Pinned against the verbatim template (oracle/ref_voc_wrap.cpp, tests/test_vocabulary.py).
void orc_bow_transform(int L, const int32_t* parent, const uint8_t* is_word, const uint8_t* ndesc, const double* nweight, int n_nodes,
                       const uint8_t* desc, int n, int levelsup, int32_t* word_id, int32_t* node_id, double* weight) {
    std::vector<std::vector<int> > children(n_nodes + 1);
    std::vector<int> word(n_nodes + 1, 0);
    int nwords = 0;
    for (int i = 1; i <= n_nodes; ++i) {
        children[parent[i - 1]].push_back(i);
        if (is_word[i - 1] > 0) word[i] = nwords++;
    }
    const int nid_level = L - levelsup;
    for (int f = 0; f < n; ++f) {
        int nid = 0;                                    // the reference leaves *nid untouched when the level is never reached
        int final_id = 0, current_level = 0;
        do {
            ++current_level;
            const std::vector<int>& nodes = children[final_id];
            final_id = nodes[0];
            double best_d = hamming(desc + (size_t)f * 32, ndesc + (size_t)(final_id - 1) * 32);
            for (size_t c = 1; c < nodes.size(); ++c) {
                const double d = hamming(desc + (size_t)f * 32, ndesc + (size_t)(nodes[c] - 1) * 32);
                if (d < best_d) { best_d = d; final_id = nodes[c]; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (!children[final_id].empty());
        word_id[f] = word[final_id]; node_id[f] = nid; weight[f] = nweight[final_id - 1];
    }
}

void orc_bruteforce_top2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* best_idx, int32_t* best_dist,
                         int32_t* second_dist) {
    for (int i = 0; i < nq; ++i) {
        int b1 = 256 + 1, b2 = 256 + 1, bi = -1;
        for (int j = 0; j < nt; ++j) {
            const int d = hamming(q + (size_t)i * 32, t + (size_t)j * 32);
            if (d < b1) { b2 = b1; b1 = d; bi = j; }
            else if (d < b2) b2 = d;
        }
        best_idx[i] = bi; best_dist[i] = b1; second_dist[i] = b2;
    }
}

}  // extern "C"

// cv::fisheye::undistortPoints(distorted, undistorted, K, D, noArray(), K) as OpenCV 4.13 evaluates it (pinned against cv2 by
// tests/golden/undistort.npz): Frame::UndistortKeyPoints / ComputeImageBounds, src/Frame.cc:638-669, 741-795.
extern "C" void orc_fisheye_undistort(const float* pts, int n, const float* K, const float* D, float* out) {
    const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const double k0 = D[0], k1 = D[1], k2 = D[2], k3 = D[3];
    const double half_pi = 3.1415926535897932384626433832795 / 2.;
    for (int i = 0; i < n; ++i) {
        const double wx = ((double)pts[2 * i] - cx) / fx, wy = ((double)pts[2 * i + 1] - cy) / fy;
        double theta_d = std::sqrt(wx * wx + wy * wy);
        theta_d = std::min(std::max(-half_pi, theta_d), half_pi);
        bool converged = false;
        double theta = theta_d, scale = 0.0;
        if (std::fabs(theta_d) > 1e-8) {
            for (int j = 0; j < 10; ++j) {
                const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
                const double a = k0 * t2, b = k1 * t4, c = k2 * t6, d = k3 * t8;
                const double fix = (theta * (1 + a + b + c + d) - theta_d) / (1 + 3 * a + 5 * b + 7 * c + 9 * d);
                theta = theta - fix;
                if (std::fabs(fix) < 1e-8) { converged = true; break; }
            }
            scale = std::tan(theta) / theta_d;
        } else {
            converged = true;
        }
        const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
        if (converged && !flipped) {
            const double ux = wx * scale, uy = wy * scale;
            out[2 * i] = (float)(fx * ux + cx);
            out[2 * i + 1] = (float)(fy * uy + cy);
        } else {
            out[2 * i] = -1000000.0f; out[2 * i + 1] = -1000000.0f;
        }
    }
}
