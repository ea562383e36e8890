// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of what the reference's bird-view feature path asks OpenCV to do (row f-3 of the scope table):
//     cv::Ptr<cv::ORB> extractorBird = cv::ORB::create(2000);
//     extractorBird->detect(mBirdviewImg, preKeysBird, mBirdviewMask);          src/Frame.cc:336-338
//     ... GuidenceKeyBirdPts, cv::cornerSubPix (oracle/bird_oracle.cpp) ...      src/Frame.cc:342-352
//     extractorBird->compute(mBirdviewImg, mvKeysBird, mDescriptorsBird);       src/Frame.cc:355
// cv::ORB is third-party code that is NOT vendored in /root/reference and no C++ OpenCV exists in the build image, so the
// algorithm is restated from OpenCV's published implementation (modules/features2d/src/orb.cpp, keypoint.cpp, fast.cpp,
// imgproc resize.cpp `INTER_LINEAR_EXACT`) for the default parameters of ORB::create(nfeatures): scaleFactor 1.2, 8 levels,
// edgeThreshold 31, firstLevel 0, WTA_K 2, HARRIS_SCORE, patchSize 31, fastThreshold 20 -- and PINNED to the only executable
// copy available, Python cv2 4.13.0: committed known answers (tests/golden/cvorb.npz, tools/gen_golden_cvorb.py) and live
// cv2 calls in tests/test_cvorb_oracle.py (keypoints incl. order, float responses / angles as bit patterns, descriptors).
//
// Two facts about the ORDER of the keypoints cv::ORB returns, both reproduced here because the reference's downstream
// indices (mvKeysBird, BirdviewMatch results) depend on it:
//   * KeyPointsFilter::retainBest is std::nth_element + std::partition on the keypoint vector: the surviving keypoints come
//     out in the order libstdc++'s introselect leaves them in.  The restatement calls the same two std:: algorithms on a
//     vector of the same element type semantics (the algorithms only compare and swap), so it follows whatever this
//     toolchain's libstdc++ does; cv2 4.13's wheel was built against a libstdc++ whose introselect is the same algorithm
//     (pinned by the order-sensitive tests).
//   * all keypoints whose response ties with the n-th best are kept (so more than n may survive).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "prim.hpp"

using namespace fbe_oracle;

namespace {

struct Kp { float x, y, size, angle, response; int32_t octave, class_id; };
static_assert(sizeof(Kp) == 28, "cv::KeyPoint layout");

static const int8_t kOrbPattern[256 * 4] = {
#include "orb_pattern.inc"
};

constexpr int kNLevels = 8, kEdge = 31, kPatch = 31, kHalfPatch = 15, kFastTh = 20, kHarrisBlock = 7;
constexpr float kHarrisK = 0.04f;

// ---- cv::resize(..., INTER_LINEAR_EXACT) for 8UC1 (imgproc/src/resize.cpp, resize_bitExact / interpolationLinear<uchar>):
// 8.8 fixed-point coefficients from double arithmetic (softdouble == IEEE double), horizontal pass in ufixedpoint16, vertical in
// ufixedpoint32, rounded half up.  Destination coordinates whose source falls left of pixel 0 / right of the last pixel copy
// the edge pixel.
struct ExactTab { std::vector<int> ofs; std::vector<int> c1; int min_ofs, max_ofs; };

ExactTab exact_tab(int src, int dst) {
    ExactTab t;
    t.ofs.assign(dst, 0); t.c1.assign(dst, 0);
    t.min_ofs = 0; t.max_ofs = dst;
    const double inv_scale = (double)dst / src;
    const double scale = 1.0 / inv_scale;
    for (int v = 0; v < dst; ++v) {
        const double fval = scale * ((double)v + 0.5) - 0.5;
        const int ival = cv_floor(fval);
        if (ival >= 0 && src > 1) {
            if (ival < src - 1) {
                t.ofs[v] = ival;
                t.c1[v] = cv_round((fval - (double)ival) * 256.0);
            } else {
                t.ofs[v] = src - 1;
                t.max_ofs = std::min(t.max_ofs, v);
            }
        } else {
            t.min_ofs = std::max(t.min_ofs, v + 1);
        }
    }
    return t;
}

void resize_linear_exact_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep) {
    const ExactTab tx = exact_tab(sw, dw), ty = exact_tab(sh, dh);
    std::vector<uint32_t> h0(dw), h1(dw);
    auto hline = [&](const uint8_t* S, std::vector<uint32_t>& H) {       // ufixedpoint16 row (value << 8)
        for (int x = 0; x < dw; ++x) {
            if (x < tx.min_ofs) H[x] = (uint32_t)S[0] << 8;
            else if (x >= tx.max_ofs) H[x] = (uint32_t)S[sw - 1] << 8;
            else {
                const int o = tx.ofs[x], c1 = tx.c1[x], c0 = 256 - c1;
                H[x] = (uint32_t)c0 * S[o] + (uint32_t)c1 * S[o + 1];
            }
        }
    };
    for (int y = 0; y < dh; ++y) {
        uint8_t* D = dst + (size_t)y * dstep;
        if (y < ty.min_ofs || y >= ty.max_ofs) {
            hline(src + (size_t)(y < ty.min_ofs ? 0 : sh - 1) * sstep, h0);
            for (int x = 0; x < dw; ++x) D[x] = (uint8_t)((h0[x] + 128u) >> 8);
        } else {
            const int o = ty.ofs[y], c1 = ty.c1[y], c0 = 256 - c1;
            hline(src + (size_t)o * sstep, h0);
            hline(src + (size_t)(o + 1) * sstep, h1);
            for (int x = 0; x < dw; ++x) D[x] = (uint8_t)(((uint32_t)c0 * h0[x] + (uint32_t)c1 * h1[x] + 32768u) >> 16);
        }
    }
}

struct Level {
    int w = 0, h = 0;
    float scale = 1.f;                 // layerScale[level] = (float)pow(1.2, level)
    std::vector<uint8_t> img;          // padded (w + 64) x (h + 64), border 32 = BORDER_REFLECT_101
    std::vector<uint8_t> mask;         // w x h or empty
    const uint8_t* at(int x, int y) const { return img.data() + (size_t)(y + 32) * (w + 64) + (x + 32); }
    uint8_t* at(int x, int y) { return img.data() + (size_t)(y + 32) * (w + 64) + (x + 32); }
    int step() const { return w + 64; }
};

// the image (and mask) pyramid of ORB_Impl::detectAndCompute: level 0 is the image, level l is resized from level l-1
// (INTER_LINEAR_EXACT); masks are resized the same way and thresholded (> 254 kept) from level 1 on
void build_pyramid(const uint8_t* img, int rows, int cols, size_t step, const uint8_t* mask, size_t mask_step, int nlevels, std::vector<Level>& L) {
    L.assign(nlevels, Level());
    for (int l = 0; l < nlevels; ++l) {
        Level& v = L[l];
        v.scale = (float)std::pow((double)1.2f, (double)l);         // getScale(): (float)std::pow(scaleFactor, (double)(level - firstLevel)), scaleFactor is a double member holding 1.2f
        const float inv = 1.0f / v.scale;
        v.w = cv_round(cols * inv); v.h = cv_round(rows * inv);
        std::vector<uint8_t> roi((size_t)v.w * v.h);
        if (l == 0) {
            for (int y = 0; y < rows; ++y) std::memcpy(roi.data() + (size_t)y * cols, img + (size_t)y * step, cols);
        } else {
            const Level& p = L[l - 1];
            resize_linear_exact_u8(p.at(0, 0), p.w, p.h, p.step(), roi.data(), v.w, v.h, v.w);
        }
        v.img.assign((size_t)(v.w + 64) * (v.h + 64), 0);
        border_reflect101_u8(roi.data(), v.w, v.h, v.w, v.img.data(), v.w + 64, 32);
        if (mask) {
            v.mask.assign((size_t)v.w * v.h, 0);
            if (l == 0) {
                // cv2 4.13 treats the mask as a predicate: any non-zero value counts as "keep" on EVERY level (observed: masks of
                // constant value 1, 7, 128 or 254 give the keypoints of the unmasked image), i.e. the mask is binarised before its
                // pyramid is built; older OpenCV releases resize the raw values and lose everything below 255 from level 1 on
                for (int y = 0; y < rows; ++y)
                    for (int x = 0; x < cols; ++x) v.mask[(size_t)y * cols + x] = mask[(size_t)y * mask_step + x] ? 255 : 0;
            } else {
                const Level& p = L[l - 1];
                resize_linear_exact_u8(p.mask.data(), p.w, p.h, p.w, v.mask.data(), v.w, v.h, v.w);
                for (uint8_t& m : v.mask) m = m > 254 ? m : 0;       // threshold(currMask, currMask, 254, 0, THRESH_TOZERO)
            }
        }
    }
}

struct ResponseGreater { bool operator()(const Kp& a, const Kp& b) const { return a.response > b.response; } };

// KeyPointsFilter::retainBest (features2d/src/keypoint.cpp)
void retain_best(std::vector<Kp>& k, int n_points) {
    if (n_points >= 0 && k.size() > (size_t)n_points) {
        if (n_points == 0) { k.clear(); return; }
        std::nth_element(k.begin(), k.begin() + n_points - 1, k.end(), ResponseGreater());
        const float ambiguous = k[n_points - 1].response;
        std::vector<Kp>::iterator new_end = std::partition(k.begin() + n_points, k.end(), [ambiguous](const Kp& p) { return p.response >= ambiguous; });
        k.resize(new_end - k.begin());
    }
}

// KeyPointsFilter::runByImageBorder: Rect(Point(b, b), Point(w - b, h - b)).contains(Point(pt)) with Point2f -> Point2i by cvRound
void run_by_image_border(std::vector<Kp>& k, int w, int h, int b) {
    if (b <= 0) return;
    if (h <= 2 * b || w <= 2 * b) { k.clear(); return; }
    std::vector<Kp> o;
    for (const Kp& p : k) {
        const int x = cv_round(p.x), y = cv_round(p.y);
        if (x >= b && x < w - b && y >= b && y < h - b) o.push_back(p);
    }
    k.swap(o);
}

void features_per_level(int nfeatures, int nlevels, std::vector<int>& per) {
    per.assign(nlevels, 0);
    const float factor = (float)(1.0 / (double)1.2f);
    float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        per[l] = cv_round(nd);
        sum += per[l];
        nd *= factor;
    }
    per[nlevels - 1] = std::max(nfeatures - sum, 0);
}

void umax_table(std::vector<int>& umax) {
    umax.assign(kHalfPatch + 2, 0);
    int v, v0, vmax = cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(kHalfPatch * std::sqrt(2.f) / 2);
    for (v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt((double)kHalfPatch * kHalfPatch - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

// HarrisResponses (orb.cpp), blockSize 7
float harris_response(const Level& L, int x0, int y0) {
    const int step = L.step(), r = kHarrisBlock / 2;
    const float scale = 1.f / ((1 << 2) * kHarrisBlock * 255.f);
    const float scale_sq_sq = scale * scale * scale * scale;
    int a = 0, b = 0, c = 0;
    for (int i = 0; i < kHarrisBlock; ++i)
        for (int j = 0; j < kHarrisBlock; ++j) {
            const uint8_t* p = L.at(x0 - r + j, y0 - r + i);
            const int Ix = (p[1] - p[-1]) * 2 + (p[-step + 1] - p[-step - 1]) + (p[step + 1] - p[step - 1]);
            const int Iy = (p[step] - p[-step]) * 2 + (p[step - 1] - p[-step - 1]) + (p[step + 1] - p[-step + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    return ((float)a * b - (float)c * c - kHarrisK * ((float)a + b) * ((float)a + b)) * scale_sq_sq;
}

// ICAngles (orb.cpp)
float ic_angle(const Level& L, int x0, int y0, const std::vector<int>& umax) {
    const int step = L.step();
    const uint8_t* center = L.at(x0, y0);
    int m_01 = 0, m_10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m_10 += u * center[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int v_sum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return fast_atan2_deg((float)m_01, (float)m_10);
}

// GaussianBlur(level ROI, 7x7, sigma 2, BORDER_REFLECT_101) as cv::ORB gets it: the ROI is a SUBMATRIX of the pyramid buffer, so
// cv::GaussianBlur skips its bit-exact 8-bit path (smooth.dispatch.cpp: only for BORDER_ISOLATED or non-submatrix sources) and
// runs sepFilter2D with the FLOAT kernel getGaussianKernel(7, 2, CV_32F): RowFilter<uchar, float> (taps accumulated left to
// right), SymmColumnFilter<Cast<float, uchar>> (centre tap, then k * (row above + row below) outwards), result cvRound-ed and
// saturated.  On every AVX2 machine OpenCV's dispatched build contracts each multiply-add to an FMA; this formulation equals
// cv2.sepFilter2D of cv2 4.13.0 bit for bit (tests/test_cvorb_oracle.py) -- it is NOT the integer blur of oracle/prim.hpp,
// which is what a standalone cv::GaussianBlur on a whole image (ORBextractor's case) computes.
void gauss7_float_u8(const Level& v, uint8_t* dst) {
    static const uint32_t kbits[4] = {0x3d8fafb1u, 0x3e06387eu, 0x3e434a39u, 0x3e5d4ae0u};     // taps 0..3 (3 = centre), float32 bit patterns
    float k[4];
    std::memcpy(k, kbits, sizeof(k));
    const float kx[7] = {k[0], k[1], k[2], k[3], k[2], k[1], k[0]};
    const int w = v.w, h = v.h;
    std::vector<float> tmp((size_t)w * (h + 6));
    for (int y = -3; y < h + 3; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t* S = v.at(x - 3, y);
            float s = kx[0] * (float)S[0];
            for (int t = 1; t < 7; ++t) s = fmaf(kx[t], (float)S[t], s);
            tmp[(size_t)(y + 3) * w + x] = s;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float* c = tmp.data() + (size_t)(y + 3) * w + x;
            float s = k[3] * c[0];
            for (int t = 1; t <= 3; ++t) s = fmaf(kx[3 + t], c[(size_t)t * w] + c[-(ptrdiff_t)t * w], s);
            const int r = cv_round(s);
            dst[(size_t)y * w + x] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
        }
}
}  // namespace

extern "C" {
void orc_resize_linear_exact_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep) {
    resize_linear_exact_u8(src, sw, sh, (size_t)sstep, dst, dw, dh, (size_t)dstep);
}

// An input on which std::nth_element(first, first + nth, last, ResponseGreater) of THIS libstdc++ runs out of its introselect depth
// budget and falls back to heap select: McIlroy's adversary ("A Killer Adversary for Quicksort", 1999) answers the comparisons
// of a real std::nth_element run so that every pivot is extreme, fixing the values as late as possible; the responses it leaves
// are distinct and consistent with every answer given, so a replay on them makes the same comparisons.  Returns the number of
// comparisons of the adversarial run (diagnostic).
int orc_antiselect(int n, int nth, float* response) {
    std::vector<int> val(n, n);                 // n = "gas": not fixed yet, compares after every fixed value
    int nsolid = 0, candidate = 0;
    long ncmp = 0;
    std::vector<int> p(n);
    for (int i = 0; i < n; ++i) p[i] = i;
    auto precedes = [&](int x, int y) {         // plays the role of ResponseGreater: true <=> x has the larger response
        ++ncmp;
        if (val[x] == n && val[y] == n) { if (x == candidate) val[x] = nsolid++; else val[y] = nsolid++; }
        if (val[x] == n) candidate = x; else if (val[y] == n) candidate = y;
        return val[x] < val[y];
    };
    if (n > 0) std::nth_element(p.begin(), p.begin() + std::min(std::max(nth, 0), n - 1), p.end(), precedes);
    for (int i = 0; i < n; ++i) if (val[i] == n) val[i] = nsolid++;
    for (int i = 0; i < n; ++i) response[i] = (float)(n - val[i]);       // earlier in the order = larger response; exact below 2^24
    return (int)std::min<long>(ncmp, 0x7fffffff);
}

// KeyPointsFilter::retainBest on a bare response array: order[n] = the permutation std::nth_element + std::partition leave behind,
// returns the number kept
int orc_retain_best(const float* response, int n, int n_points, int32_t* order) {
    std::vector<Kp> k(n);
    for (int i = 0; i < n; ++i) { k[i] = Kp{0, 0, 0, 0, response[i], 0, i}; }
    const size_t before = k.size();
    std::vector<Kp> all = k;
    if (n_points >= 0 && k.size() > (size_t)n_points && n_points > 0) {
        std::nth_element(all.begin(), all.begin() + n_points - 1, all.end(), ResponseGreater());
        const float ambiguous = all[n_points - 1].response;
        std::vector<Kp>::iterator new_end = std::partition(all.begin() + n_points, all.end(), [ambiguous](const Kp& p) { return p.response >= ambiguous; });
        for (size_t i = 0; i < before; ++i) order[i] = all[i].class_id;
        return (int)(new_end - all.begin());
    }
    for (size_t i = 0; i < before; ++i) order[i] = (int)i;
    return n_points == 0 ? 0 : n;
}

// the blur of one pyramid level as cv::ORB computes it (see gauss7_float_u8), on a stand-alone image with REFLECT_101 borders
void orc_cvorb_blur(const uint8_t* img, int rows, int cols, int step, uint8_t* dst) {
    Level v;
    v.w = cols; v.h = rows;
    std::vector<uint8_t> roi((size_t)cols * rows);
    for (int y = 0; y < rows; ++y) std::memcpy(roi.data() + (size_t)y * cols, img + (size_t)y * step, cols);
    v.img.assign((size_t)(cols + 64) * (rows + 64), 0);
    border_reflect101_u8(roi.data(), cols, rows, cols, v.img.data(), cols + 64, 32);
    gauss7_float_u8(v, dst);
}

// cv::ORB::create(nfeatures)->detect(img, keypoints, mask); mask may be NULL.  Returns the number of keypoints (in cv::ORB's
// output order); at most `cap` are written.  counts (may be NULL): [0] FAST corners before any filter, summed over levels.
int orc_cvorb_detect(const uint8_t* img, int rows, int cols, int step, const uint8_t* mask, int mask_step, int nfeatures, Kp* out, int cap,
                     int32_t* counts) {
    std::vector<Level> L;
    build_pyramid(img, rows, cols, (size_t)step, mask, (size_t)mask_step, kNLevels, L);
    std::vector<int> per, umax;
    features_per_level(nfeatures, kNLevels, per);
    umax_table(umax);
    std::vector<Kp> all;
    std::vector<int> counters(kNLevels, 0);
    int nfast = 0;
    for (int l = 0; l < kNLevels; ++l) {
        const Level& v = L[l];
        std::vector<FastKp> f;
        fast9_nms(v.at(0, 0), v.w, v.h, v.step(), kFastTh, f);          // FastFeatureDetector::create(20, true)->detect(img, keypoints, mask)
        nfast += (int)f.size();
        std::vector<Kp> k;
        for (const FastKp& p : f) {
            if (!v.mask.empty() && v.mask[(size_t)(int)(p.y + 0.5f) * v.w + (int)(p.x + 0.5f)] == 0) continue;   // runByPixelsMask
            k.push_back(Kp{(float)p.x, (float)p.y, 7.f, -1.f, (float)p.score, 0, -1});
        }
        run_by_image_border(k, v.w, v.h, kEdge);
        retain_best(k, 2 * per[l]);                                      // HARRIS_SCORE: keep twice as many by FAST score first
        counters[l] = (int)k.size();
        for (Kp& p : k) { p.octave = l; p.size = kPatch * v.scale; }
        all.insert(all.end(), k.begin(), k.end());
    }
    if (counts) counts[0] = nfast;
    if (all.empty()) return 0;
    for (Kp& p : all) p.response = harris_response(L[p.octave], cv_round(p.x), cv_round(p.y));
    std::vector<Kp> sel;
    int offset = 0;
    for (int l = 0; l < kNLevels; ++l) {
        std::vector<Kp> k(all.begin() + offset, all.begin() + offset + counters[l]);
        offset += counters[l];
        retain_best(k, per[l]);
        sel.insert(sel.end(), k.begin(), k.end());
    }
    for (Kp& p : sel) p.angle = ic_angle(L[p.octave], cv_round(p.x), cv_round(p.y), umax);
    for (Kp& p : sel) { const float s = L[p.octave].scale; p.x *= s; p.y *= s; }
    const int n = (int)sel.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = sel[i];
    return n;
}

// cv::ORB::create(...)->compute(img, keypoints, descriptors): keypoints within 31 px of the image border are removed, keypoints
// are regrouped by octave when they are not sorted by it, every level present (0 .. max octave) is blurred, then one 256-bit
// steered-BRIEF descriptor per keypoint.  kps is rewritten with the surviving keypoints; returns their number.
int orc_cvorb_compute(const uint8_t* img, int rows, int cols, int step, Kp* kps, int n, uint8_t* desc) {
    if (n <= 0) return 0;
    int nlevels = 0;
    bool sorted = true;
    for (int i = 0; i < n; ++i) {
        if (kps[i].octave < 0) return -1;
        if (i > 0 && kps[i].octave < kps[i - 1].octave) sorted = false;
        nlevels = std::max(nlevels, (int)kps[i].octave);
    }
    nlevels++;
    std::vector<Level> L;
    build_pyramid(img, rows, cols, (size_t)step, nullptr, 0, nlevels, L);
    std::vector<Kp> k(kps, kps + n);
    run_by_image_border(k, cols, rows, kEdge);
    if (!sorted) {
        std::vector<std::vector<Kp> > by(nlevels);
        for (const Kp& p : k) by[p.octave].push_back(p);
        k.clear();
        for (int l = 0; l < nlevels; ++l) k.insert(k.end(), by[l].begin(), by[l].end());
    }
    if (k.empty()) return 0;
    for (int l = 0; l < nlevels; ++l) {                               // GaussianBlur(workingMat, workingMat, Size(7, 7), 2, 2, BORDER_REFLECT_101) on the level ROI
        Level& v = L[l];
        std::vector<uint8_t> b((size_t)v.w * v.h);
        gauss7_float_u8(v, b.data());
        for (int y = 0; y < v.h; ++y) std::memcpy(v.at(0, y), b.data() + (size_t)y * v.w, v.w);
    }
    for (size_t j = 0; j < k.size(); ++j) {
        const Kp& p = k[j];
        const Level& v = L[p.octave];
        const float scale = 1.f / v.scale;
        float angle = p.angle;
        angle *= (float)(3.14159265358979323846 / 180.f);
        // orb.cpp writes (float)cos(angle) on a float: evaluated here as the correctly rounded cosine (double cos, rounded once).
        // cosf, FMA-contracted and plain forms of the rotation below were compared against cv2 4.13 on 70 662 keypoints with
        // random angles (36 M rotated samples): all give cv2's descriptors, so the form the GPU can reproduce exactly is used.
        const float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
        const int step_ = v.step();
        const uint8_t* center = v.at(cv_round(p.x * scale), cv_round(p.y * scale));
        uint8_t* d = desc + j * 32;
        const int8_t* pat = kOrbPattern;
        for (int i = 0; i < 32; ++i, pat += 32) {
            int val = 0;
            for (int bit = 0; bit < 8; ++bit) {
                const int8_t* q = pat + 4 * bit;
                float x = q[0] * a - q[1] * b, y = q[0] * b + q[1] * a;
                const int t0 = center[cv_round(y) * step_ + cv_round(x)];
                x = q[2] * a - q[3] * b; y = q[2] * b + q[3] * a;
                const int t1 = center[cv_round(y) * step_ + cv_round(x)];
                val |= (t0 < t1) << bit;
            }
            d[i] = (uint8_t)val;
        }
    }
    for (size_t j = 0; j < k.size(); ++j) kps[j] = k[j];
    return (int)k.size();
}

}  // extern "C"
