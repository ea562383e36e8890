// Orientation (IC_Angle, src/ORBextractor.cc:77-104), rotated BRIEF-256 (computeOrbDescriptor :108-147) and the
// operator() epilogue (:1059-1104: octave/size fix-up :837-847, pt *= scale for level > 0, level concatenation).
// One CTA = 32 keypoints, 8 warps x 4 keypoints:
//   A. moments: lanes are COLUMNS u = -15..15 of the radius-15 disc of the UNBLURRED level and the warp walks the 31 rows,
//      so every load is one coalesced 31-byte row segment (int32 sums, exact); warp-reduced;
//      angle = fastAtan2(m01, m10) -- OpenCV's 7th-order polynomial, evaluated with explicit round-to-nearest fp32
//      mul/add/div (no FMA contraction) in the reference's operation order.
//   B. a = cos, b = sin of angle*(float)(CV_PI/180.f), computed in double and rounded (the reference calls float cos/sin
//      from libm, which are correctly rounded for all but a ~1e-3 fraction of arguments): ONE THREAD per keypoint, so the
//      fp64 pipe sees 1/32 of the warp-instructions a warp-per-keypoint evaluation would issue.
//   C. descriptor: the warp stages the 37x37 patch of the BLURRED level into shared memory with aligned 32-bit loads
//      (pattern radius 13 -> rotated reach <= 18), then lane = output byte, 16 rotated samples each, read from shared memory;
//      cvRound(x*b + y*a), cvRound(x*a - y*b) as __float2int_rn of un-contracted fp32 products.
#include <cuda_fp16.h>
#include "fbe_internal.cuh"
#include "orb_device.cuh"

namespace fbe {

__constant__ int8_t c_pattern[256 * 4] = {
#include "orb_pattern.inc"
};

constexpr int kDescWarps = 8;        // keypoints per warp (kDescPerWarp) is a template parameter: 4 for batches, 1 when a call is latency-bound
constexpr int kCoefRows = 37;         // 11 passes x 3 rows + the 5 idle lanes' reach, rows >= 31 hold zeros
constexpr int kPatchR = 18, kPatchRows = 2 * kPatchR + 1, kPatchWords = 12;      // 37 rows x 44 bytes, stored with a pitch of 12 words:
                                                                                 // the staging stores of 8 rows x 4 words then hit 32 distinct banks (pitch 11: 2-3 way conflicts)

template <int kDescPerWarp>
__global__ void __launch_bounds__(kDescWarps * 32, 8) k_describe(const Plan* __restrict__ plan, Workspace ws) {
    constexpr int kDescPerCta = kDescWarps * kDescPerWarp;
    __shared__ uint2 s_pat[256];                                        // [j][lane]: pair lane*8 + j as (x0, y0, x1, y1) in fp16 (|coordinate| <= 13: exact):
                                                                        // 8 bytes per lane = 2 shared-memory wavefronts per load instead of 4
    __shared__ uint32_t s_patch[kDescWarps][kPatchRows * kPatchWords];
    __shared__ uint32_t s_key[kDescPerCta];
    __shared__ int s_lvl[kDescPerCta];
    __shared__ float s_angle[kDescPerCta], s_cos[kDescPerCta], s_sin[kDescPerCta];
    __shared__ unsigned s_coef[2 * kCoefRows * 9];

    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.y;
    const int* level_n = ws.level_n + (size_t)b * FBE_MAX_LEVELS;
    const int nl = plan->nlevels;
    int total = 0;
    for (int i = 0; i < nl; ++i) total += level_n[i];
    if (blockIdx.x == 0 && tid == 0) ws.out_n[b] = total;
    const int g0 = blockIdx.x * kDescPerCta;
    if (g0 >= total) return;
    {
        const int p = (tid & 31) * 8 + (tid >> 5);                       // s_pat[tid] = pair p
        const __half2 h01 = __floats2half2_rn((float)c_pattern[4 * p], (float)c_pattern[4 * p + 1]);
        const __half2 h23 = __floats2half2_rn((float)c_pattern[4 * p + 2], (float)c_pattern[4 * p + 3]);
        s_pat[tid] = make_uint2(*reinterpret_cast<const unsigned*>(&h01), *reinterpret_cast<const unsigned*>(&h23));
    }

    // ---- A: IC_Angle on the unblurred level ---------------------------------------------------------------------
    // A disc row is 31 bytes = 9 aligned words; lanes 0..26 hold (row group 0..2, word 0..8), so one coalesced load fetches three
    // rows and 11 passes cover the 31 rows.  Neighbouring lanes exchange words (SHFL) to realign them to the window u = -15 .. 16
    // (funnel shift), and the row sums are integer dot products (IDP.4A) with coefficient words from shared memory: sum(I) with
    // the disc mask, sum((u + 15) * I) with the masked weights; m10 = sum((u + 15) I) - 15 sum(I), m01 = sum(v * rowsum) -- exact ints.
    {
        // s_coef[r * 9 + i]: mask word i of row r; s_coef[kCoefRows * 9 + r * 9 + i]: weight word (rows >= 31 and word 8 are zero)
        for (int k = tid; k < 2 * kCoefRows * 9; k += kDescWarps * 32) {
            const int kk = k % (kCoefRows * 9), r = kk / 9, i = kk - 9 * r, av = r < 15 ? 15 - r : r - 15;
            unsigned wv = 0;
            if (r < 31 && i < 8) {
                const int d = kUmaxDev(av);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int u = -15 + 4 * i + j;
                    if (u >= -d && u <= d) wv |= (k < kCoefRows * 9 ? 1u : (unsigned)(u + 15)) << (8 * j);
                }
            }
            s_coef[k] = wv;
        }
    }
    __syncthreads();
    const int grp = lane / 9, wi = lane - 9 * grp;
#pragma unroll 1
    for (int k = 0; k < kDescPerWarp; ++k) {
        const int slot = wid * kDescPerWarp + k, gidx = g0 + slot;
        if (gidx >= total) break;
        int l = 0, off = 0;
        while (gidx >= off + level_n[l]) { off += level_n[l]; ++l; }
        const LevelGeom& g = plan->lv[l];
        const uint32_t key = ws.sel[(size_t)b * plan->kp_cap_total + g.kp_base + (gidx - off)];
        const int kx = key_x(key), ky = key_y(key), pw = g.pitch >> 2;              // pitch is a multiple of 16: every row shares the alignment
        const size_t byte0 = (size_t)b * plan->pyr_bytes + g.img_off + (size_t)(ky + kEdge - 15) * g.pitch + (kx + kEdge - 15);     // (u, v) = (-15, -15)
        const int sh = (int)(byte0 & 3) * 8;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(ws.pyr + (byte0 & ~(size_t)3)) + wi + grp * pw;
        unsigned RS = 0, US = 0;
        int T = 0;
#pragma unroll
        for (int p = 0; p < 11; ++p) {
            const int r = 3 * p + grp;
            const uint32_t x = (lane < 27 && r < 31) ? __ldg(w + 3 * p * pw) : 0u;
            const uint32_t xn = __shfl_down_sync(0xffffffffu, x, 1);
            const unsigned q = __funnelshift_r(x, xn, sh);
            const unsigned a = __dp4a(q, s_coef[27 * p + lane], 0u);             // lanes >= 27 read zero rows of the table
            RS += a;
            T += (r - 15) * (int)a;
            US = __dp4a(q, s_coef[kCoefRows * 9 + 27 * p + lane], US);
        }
        int m10 = (int)US - 15 * (int)RS, m01 = T;
        m10 = __reduce_add_sync(0xffffffffu, m10);
        m01 = __reduce_add_sync(0xffffffffu, m01);
        if (lane == 0) {
            s_angle[slot] = fast_atan2_deg((float)m01, (float)m10);
            s_key[slot] = key;
            s_lvl[slot] = l;
        }
    }
    __syncthreads();

    // ---- B: rotation, one thread per keypoint ----------------------------------------------------------------------
    if (tid < kDescPerCta && g0 + tid < total) {
        const float factorPI = (float)(3.14159265358979323846 / 180.0);    // (float)(CV_PI/180.f)
        const float ang = __fmul_rn(s_angle[tid], factorPI);
        s_cos[tid] = (float)cos((double)ang);
        s_sin[tid] = (float)sin((double)ang);
    }
    __syncthreads();

    // ---- C: rotated BRIEF on the blurred level ---------------------------------------------------------------------
    uint32_t* patch = s_patch[wid];
    const uint8_t* patch8 = reinterpret_cast<const uint8_t*>(patch);
#pragma unroll 1
    for (int k = 0; k < kDescPerWarp; ++k) {
        const int slot = wid * kDescPerWarp + k, gidx = g0 + slot;
        if (gidx >= total) break;
        const uint32_t key = s_key[slot];
        const int l = s_lvl[slot];
        const LevelGeom& g = plan->lv[l];
        const int kx = key_x(key), ky = key_y(key), pitch = g.pitch;
        const int pcol = kx + kEdge - kPatchR;                    // padded column of the patch's left edge
        const int shift = pcol & 3;
        const uint8_t* src = ws.blur + (size_t)b * plan->pyr_bytes + g.img_off + (size_t)(ky + kEdge - kPatchR) * pitch + (pcol - shift);
        __syncwarp();
        {   // 8 rows per pass: lane = (row in pass, word quarter); words wq, wq+4, wq+8 of the 11-word row
            const int r8 = lane >> 2, wq = lane & 3;
            const uint32_t* sp = reinterpret_cast<const uint32_t*>(src + (size_t)r8 * pitch) + wq;
            uint32_t* dp = patch + r8 * kPatchWords + wq;
            const int pw8 = pitch * 2;                               // 8 rows in 32-bit words
#pragma unroll
            for (int pass = 0; pass < 5; ++pass) {
                if (pass * 8 + r8 < kPatchRows) {
                    dp[0] = __ldg(sp);
                    dp[4] = __ldg(sp + 4);
                    if (wq < 3) dp[8] = __ldg(sp + 8);
                }
                sp += pw8;
                dp += 8 * kPatchWords;
            }
        }
        __syncwarp();
        const float a = s_cos[slot], bb = s_sin[slot];
        const uint8_t* pc = patch8 + kPatchR * (kPatchWords * 4) + kPatchR + shift;      // patch centre
        unsigned val = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint2 ph = s_pat[j * 32 + lane];
            const float2 p01 = __half22float2(*reinterpret_cast<const __half2*>(&ph.x)), p23 = __half22float2(*reinterpret_cast<const __half2*>(&ph.y));
            const float4 p = make_float4(p01.x, p01.y, p23.x, p23.y);
            // cvRound of a value below 2^22 in magnitude: adding 1.5 * 2^23 rounds to the nearest integer, ties to even, exactly
            // like cvtss2si / F2I.RN, and leaves the integer in the low mantissa bits (FADD + IADD instead of the quarter-rate F2I)
            const int r0 = cv_round_small(__fadd_rn(__fmul_rn(p.x, bb), __fmul_rn(p.y, a)));
            const int c0 = cv_round_small(__fsub_rn(__fmul_rn(p.x, a), __fmul_rn(p.y, bb)));
            const int r1 = cv_round_small(__fadd_rn(__fmul_rn(p.z, bb), __fmul_rn(p.w, a)));
            const int c1 = cv_round_small(__fsub_rn(__fmul_rn(p.z, a), __fmul_rn(p.w, bb)));
            const int t0 = pc[r0 * (kPatchWords * 4) + c0], t1 = pc[r1 * (kPatchWords * 4) + c1];
            val |= (unsigned)(t0 < t1) << j;
        }
        ws.out_desc[((size_t)b * plan->kp_cap_total + gidx) * 32 + lane] = (uint8_t)val;

        // ---- keypoint record -----------------------------------------------------------------------------------
        if (lane == 0) {
            fbe_keypoint kp;
            float fx = (float)kx, fy = (float)ky;
            if (l != 0) { fx = __fmul_rn(fx, g.scale); fy = __fmul_rn(fy, g.scale); }
            kp.x = fx; kp.y = fy; kp.size = g.patch_size; kp.angle = s_angle[slot]; kp.response = (float)key_s(key);
            kp.octave = l; kp.class_id = -1;
            ws.out_kps[(size_t)b * plan->kp_cap_total + gidx] = kp;
        }
    }
}

int launch_describe(const Plan& hp, const Plan* dp, const Workspace& ws, int nimg, cudaStream_t st) {
    // a few frames cannot fill the GPU with 32 keypoints per CTA: one keypoint per warp shortens the call's critical path
    if ((size_t)nimg * hp.kp_cap_total <= (size_t)4 * 148 * kDescWarps) {
        dim3 grid((hp.kp_cap_total + kDescWarps - 1) / kDescWarps, nimg);
        FBE_CUDA(launch_dep(k_describe<1>, grid, dim3(kDescWarps * 32), 0, st, dp, ws));
    } else {
        dim3 grid((hp.kp_cap_total + 4 * kDescWarps - 1) / (4 * kDescWarps), nimg);
        FBE_CUDA(launch_dep(k_describe<4>, grid, dim3(kDescWarps * 32), 0, st, dp, ws));
    }
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
