// Device helpers shared by the two ORB flavours of this library (ORBextractor: describe.cu; cv::ORB of the bird view:
// bird_orb.cu): OpenCV's fastAtan2 and the row half-widths of the radius-15 intensity-centroid disc.
#pragma once
#include <cuda_runtime.h>

namespace fbe {

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = 2.220446049250313e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// cvRound(x) for |x| < 2^22 without the conversion pipe: (x + 1.5 * 2^23) has the rounded integer in its low mantissa bits
__device__ __forceinline__ int cv_round_small(float x) { return __float_as_int(__fadd_rn(x, 12582912.f)) - 0x4B400000; }

// row half-widths of the radius-15 disc (umax, src/ORBextractor.cc:452-469; HALF_PATCH_SIZE is a compile-time constant)
constexpr int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
// the same table for a run-time row index (a switch the compiler turns into a select chain / constant lookup)
__device__ __forceinline__ int kUmaxDev(int av) {
    return (int)((0x3689ABCDDEEEFFFFull >> (4 * av)) & 15ull);      // kUmax packed 4 bits per row, low nibble = row 0
}
template <int V> struct UmaxOf { static constexpr int value = kUmax[V < 0 ? -V : V]; };

}  // namespace fbe
