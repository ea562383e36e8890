// ORACLE (test infrastructure). C exports of the scalar primitives in prim.hpp so that
// tests/test_oracle_prims.py can pin each one against cv2 4.13.0 / committed fixtures.
#include "prim.hpp"
using namespace fbe_oracle;

extern "C" {

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw, int dh, int dstep) {
    resize_linear_u8(src, sw, sh, (size_t)sstep, dst, dw, dh, (size_t)dstep);
}

void orc_border_reflect101_u8(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep, int b) {
    border_reflect101_u8(src, w, h, (size_t)sstep, dst, (size_t)dstep, b);
}

// returns count; writes up to cap (x,y,score) triples
int orc_fast9_nms(const uint8_t* img, int w, int h, int step, int th, int32_t* xys, int cap) {
    std::vector<FastKp> v;
    fast9_nms(img, w, h, (size_t)step, th, v);
    int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) { xys[3 * i] = v[i].x; xys[3 * i + 1] = v[i].y; xys[3 * i + 2] = v[i].score; }
    return n;
}

void orc_gauss7_u8(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep) {
    gauss7_u8(src, w, h, (size_t)sstep, dst, (size_t)dstep);
}

float orc_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
int orc_cv_round_f(float v) { return cv_round(v); }
int orc_cv_round_d(double v) { return cv_round(v); }

}  // extern "C"
