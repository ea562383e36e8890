// ORACLE (test infrastructure, never shipped, never on the product path).
//
// Scalar CPU restatements of the OpenCV primitives the reference's ORB front end calls.
// OpenCV is NOT vendored in /root/reference and no C++ OpenCV exists in the build image, so
// these are restated from the published algorithms and pinned against the only executable
// copy available -- Python cv2 4.13.0 -- by tests/test_oracle_prims.py and by the committed
// fixtures under tests/golden/ (generated with cv2 by tools/gen_golden.py).
//
// Reference call sites these stand in for (file:line relative to /root/reference):
//   cv::resize INTER_LINEAR          src/ORBextractor.cc:1120
//   cv::copyMakeBorder REFLECT_101   src/ORBextractor.cc:1122-1123,1127-1128
//   cv::FAST(roi, kps, th, true)     src/ORBextractor.cc:809-810,814-815
//   cv::GaussianBlur 7x7 sigma 2     src/ORBextractor.cc:1086
//   cv::fastAtan2                    src/ORBextractor.cc:103
//   cvRound / cvFloor / cvCeil       src/ORBextractor.cc:81,115,119-120,442,456-460,1112
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

namespace fbe_oracle {

// cvRound: round-half-to-even (SSE cvtss2si / lrint under the default rounding mode).
static inline int cv_round(double v) { return (int)std::nearbyint(v); }
static inline int cv_round(float v) { return (int)std::nearbyintf(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

static inline int reflect101(int p, int len) {
    // BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

// ---- resize, INTER_LINEAR, 8UC1 (OpenCV >= 3: 11-bit fixed point, separable) -------------
struct LinearTab {
    std::vector<int> ofs;
    std::vector<short> w;   // 2 per destination coordinate
};

static inline LinearTab linear_tab(int src_len, int dst_len) {
    LinearTab t;
    t.ofs.resize(dst_len);
    t.w.resize(2 * (size_t)dst_len);
    const double scale = (double)src_len / dst_len;
    for (int d = 0; d < dst_len; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src_len - 1) { s = src_len - 1; f = 0.f; }
        t.ofs[d] = s;
        float c0 = 1.f - f, c1 = f;
        int w0 = cv_round(c0 * 2048.f), w1 = cv_round(c1 * 2048.f);
        t.w[2 * d] = (short)std::min(std::max(w0, -32768), 32767);
        t.w[2 * d + 1] = (short)std::min(std::max(w1, -32768), 32767);
    }
    return t;
}

static inline void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep,
                                    uint8_t* dst, int dw, int dh, size_t dstep) {
    LinearTab tx = linear_tab(sw, dw), ty = linear_tab(sh, dh);
    std::vector<int> r0(dw), r1(dw);
    for (int y = 0; y < dh; ++y) {
        int sy = ty.ofs[y];
        int sy1 = std::min(sy + 1, sh - 1);
        const uint8_t* S0 = src + (size_t)sy * sstep;
        const uint8_t* S1 = src + (size_t)sy1 * sstep;
        for (int x = 0; x < dw; ++x) {
            int sx = tx.ofs[x];
            int sx1 = std::min(sx + 1, sw - 1);
            int a0 = tx.w[2 * x], a1 = tx.w[2 * x + 1];
            r0[x] = S0[sx] * a0 + S0[sx1] * a1;
            r1[x] = S1[sx] * a0 + S1[sx1] * a1;
        }
        int b0 = ty.w[2 * y], b1 = ty.w[2 * y + 1];
        uint8_t* D = dst + (size_t)y * dstep;
        for (int x = 0; x < dw; ++x) {
            int v = (((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2;
            D[x] = (uint8_t)std::min(std::max(v, 0), 255);
        }
    }
}

// ---- copyMakeBorder, REFLECT_101 (isolated) ----------------------------------------------
// dst is (h+2b) x (w+2b); src may alias the interior of dst (in-place fill of the frame).
static inline void border_reflect101_u8(const uint8_t* src, int w, int h, size_t sstep,
                                        uint8_t* dst, size_t dstep, int b) {
    for (int y = 0; y < h; ++y) {
        uint8_t* D = dst + (size_t)(y + b) * dstep;
        const uint8_t* S = src + (size_t)y * sstep;
        if (D + b != S) std::memmove(D + b, S, w);
        for (int x = 0; x < b; ++x) {
            D[x] = D[b + reflect101(x - b, w)];
            D[b + w + x] = D[b + reflect101(w + x, w)];
        }
    }
    for (int y = 0; y < b; ++y) {
        std::memcpy(dst + (size_t)y * dstep, dst + (size_t)(b + reflect101(y - b, h)) * dstep, w + 2 * b);
        std::memcpy(dst + (size_t)(b + h + y) * dstep, dst + (size_t)(b + reflect101(h + y, h)) * dstep, w + 2 * b);
    }
}

// ---- FAST-9/16 with non-max suppression ---------------------------------------------------
struct FastKp { int x, y, score; };

static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// m(p) = max over the 16 arcs of 9 contiguous ring pixels of min(|I - c|) with one sign.
// p is a FAST-9 corner at threshold t  <=>  m > t ;  OpenCV's cornerScore = m - 1.
static inline int fast9_m(const uint8_t* c, size_t step) {
    int d[25];
    for (int k = 0; k < 16; ++k) d[k] = (int)c[(ptrdiff_t)kRingDy[k] * (ptrdiff_t)step + kRingDx[k]] - (int)c[0];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = 0;
    for (int s = 0; s < 16; ++s) {
        int lo = d[s], hi = d[s];
        for (int i = 1; i < 9; ++i) { lo = std::min(lo, d[s + i]); hi = std::max(hi, d[s + i]); }
        best = std::max(best, std::max(lo, -hi));
    }
    return best;
}

// Any arc of 9 contiguous ring pixels contains k or k+8 for every k, so a corner at threshold th
// needs, for one sign, max(d[k], d[k+8]) > th for all k in 0..7.  Pure early-out; never changes m.
static inline bool fast9_maybe(const uint8_t* c, size_t step, int th) {
    bool br = true, dk = true;
    for (int k = 0; k < 8 && (br || dk); ++k) {
        int a = (int)c[(ptrdiff_t)kRingDy[k] * (ptrdiff_t)step + kRingDx[k]] - (int)c[0];
        int b = (int)c[(ptrdiff_t)kRingDy[k + 8] * (ptrdiff_t)step + kRingDx[k + 8]] - (int)c[0];
        br = br && (a > th || b > th);
        dk = dk && (a < -th || b < -th);
    }
    return br || dk;
}

// cv::FAST(img, kps, th, nonmaxSuppression=true), TYPE_9_16.  Output order row-major.
static inline void fast9_nms(const uint8_t* img, int w, int h, size_t step, int th,
                             std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    // mm holds m for corners (m > th >= 0, so m >= 1) and 0 elsewhere; score = m - 1.
    std::vector<int> mm((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* c = img + (size_t)y * step + x;
            if (!fast9_maybe(c, step, th)) continue;          // exact necessary condition, speed only
            int m = fast9_m(c, step);
            if (m > th) mm[(size_t)y * w + x] = m;
        }
    auto score = [&](int x, int y) { int m = mm[(size_t)y * w + x]; return m > 0 ? m - 1 : 0; };
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            if (mm[(size_t)y * w + x] == 0) continue;
            int s = score(x, y);
            bool keep = true;
            for (int dy = -1; dy <= 1 && keep; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy) continue;
                    if (score(x + dx, y + dy) >= s) { keep = false; break; }   // strict > required
                }
            if (keep) out.push_back({x, y, s});
        }
}

// ---- GaussianBlur 7x7, sigma 2, REFLECT_101, 8U (OpenCV >= 3.4.2 bit-exact path) ----------
static const int kGauss7[7] = {18, 34, 48, 56, 48, 34, 18};   // sum 256

static inline void gauss7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep) {
    std::vector<uint32_t> tmp((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = src + (size_t)y * sstep;
        for (int x = 0; x < w; ++x) {
            uint32_t a = 0;
            for (int k = -3; k <= 3; ++k) a += kGauss7[k + 3] * S[reflect101(x + k, w)];
            tmp[(size_t)y * w + x] = a;
        }
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint32_t a = 0;
            for (int k = -3; k <= 3; ++k) a += kGauss7[k + 3] * tmp[(size_t)reflect101(y + k, h) * w + x];
            dst[(size_t)y * dstep + x] = (uint8_t)((a + 32768u) >> 16);
        }
}

// ---- fastAtan2 (degrees, [0,360)) ----------------------------------------------------------
// 7th-order odd minimax polynomial used by OpenCV >= 3; evaluated in fp32 without FMA
// contraction (this translation unit must be built with -ffp-contract=off).
static inline float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = 2.220446049250313e-16f;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

}  // namespace fbe_oracle
