// Orientation (IC_Angle, src/ORBextractor.cc:77-104), rotated BRIEF-256 (computeOrbDescriptor :108-147) and the
// operator() epilogue (:1059-1104: octave/size fix-up :837-847, pt *= scale for level > 0, level concatenation).
// One warp per keypoint:
//   * moments: lane v+15 sums row v of the radius-15 disc of the UNBLURRED level (int32, exact), warp-reduced;
//     angle = fastAtan2(m01, m10) -- OpenCV's 7th-order polynomial, evaluated with explicit round-to-nearest fp32
//     mul/add/div (no FMA contraction) in the reference's operation order.
//   * descriptor: lane = output byte; 16 rotated samples of the BLURRED level each.  The rotation uses
//     a = cos, b = sin of angle*(float)(CV_PI/180.f) rounded from double (the reference calls float cos/sin from
//     libm, which are correctly rounded for all but a ~1e-3 fraction of arguments) and
//     cvRound(x*b + y*a), cvRound(x*a - y*b) as __float2int_rn of un-contracted fp32 products.
#include "fbe_internal.cuh"

namespace fbe {

__constant__ int8_t c_pattern[256 * 4] = {
#include "orb_pattern.inc"
};

constexpr int kDescWarps = 8;

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

__global__ void __launch_bounds__(kDescWarps * 32) k_describe(const Plan* __restrict__ plan, Workspace ws) {
    __shared__ int8_t s_pat[1024];
    for (int i = threadIdx.x; i < 1024; i += kDescWarps * 32) s_pat[i] = c_pattern[i];
    __syncthreads();
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int gidx = blockIdx.x * kDescWarps + (threadIdx.x >> 5);     // output keypoint index within the image
    // level of this output index: levels are concatenated 0..L-1 in order
    const int* level_n = ws.level_n + (size_t)b * FBE_MAX_LEVELS;
    int l = 0, off = 0, total = 0;
    const int nl = plan->nlevels;
    for (int i = 0; i < nl; ++i) total += level_n[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.out_n[b] = total;
    if (gidx >= total) return;
    while (gidx >= off + level_n[l]) { off += level_n[l]; ++l; }
    const LevelGeom g = plan->lv[l];
    const uint32_t key = ws.sel[(size_t)b * plan->kp_cap_total + g.kp_base + (gidx - off)];
    const int kx = key_x(key), ky = key_y(key);

    // ---- IC_Angle on the unblurred level ---------------------------------------------------------------------
    const uint8_t* img = ws.pyr + (size_t)b * plan->pyr_bytes + g.img_off;
    const uint8_t* c = img + (size_t)(ky + kEdge) * g.pitch + (kx + kEdge);
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - 15;
        const int d = plan->umax[v < 0 ? -v : v];
        const uint8_t* row = c + (ptrdiff_t)v * g.pitch;
        int rs = 0, ws_ = 0;
        for (int u = -d; u <= d; ++u) { const int p = row[u]; rs += p; ws_ += u * p; }
        m10 = ws_;
        m01 = v * rs;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- rotated BRIEF on the blurred level ------------------------------------------------------------------
    const float factorPI = (float)(3.14159265358979323846 / 180.0);    // (float)(CV_PI/180.f)
    const float ang = __fmul_rn(angle, factorPI);
    const float a = (float)cos((double)ang), bb = (float)sin((double)ang);
    const uint8_t* bimg = ws.blur + (size_t)b * plan->pyr_bytes + g.img_off;
    const uint8_t* bc = bimg + (size_t)(ky + kEdge) * g.pitch + (kx + kEdge);
    unsigned val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int8_t* p = s_pat + (lane * 8 + j) * 4;
        const float x0 = (float)p[0], y0 = (float)p[1], x1 = (float)p[2], y1 = (float)p[3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, bb), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, bb)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, bb), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, bb)));
        const int t0 = bc[(ptrdiff_t)r0 * g.pitch + c0], t1 = bc[(ptrdiff_t)r1 * g.pitch + c1];
        val |= (unsigned)(t0 < t1) << j;
    }
    uint8_t* desc = ws.out_desc + ((size_t)b * plan->kp_cap_total + gidx) * 32;
    desc[lane] = (uint8_t)val;

    // ---- keypoint record ---------------------------------------------------------------------------------------
    if (lane == 0) {
        fbe_keypoint kp;
        float fx = (float)kx, fy = (float)ky;
        if (l != 0) { fx = __fmul_rn(fx, g.scale); fy = __fmul_rn(fy, g.scale); }
        kp.x = fx; kp.y = fy; kp.size = g.patch_size; kp.angle = angle; kp.response = (float)key_s(key);
        kp.octave = l; kp.class_id = -1;
        ws.out_kps[(size_t)b * plan->kp_cap_total + gidx] = kp;
    }
}

int launch_describe(const Plan& hp, const Plan* dp, const Workspace& ws, int nimg, cudaStream_t st) {
    dim3 grid((hp.kp_cap_total + kDescWarps - 1) / kDescWarps, nimg);
    k_describe<<<grid, kDescWarps * 32, 0, st>>>(dp, ws);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
