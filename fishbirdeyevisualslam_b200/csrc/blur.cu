// GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of every pyramid level (src/ORBextractor.cc:1085-1086), all levels of
// all images of the batch in ONE launch.
// OpenCV >= 3.4.2 evaluates this 8-bit case exactly: integer kernel [18,34,48,56,48,34,18]/256 per axis, no
// intermediate rounding, out = (sum_y sum_x k_y k_x p + 32768) >> 16 (pinned against cv2 4.13).
// The reference blurs a clone of the un-padded level with REFLECT_101 at its edge; our padded level already
// carries that reflection in its 19-px frame, so the stencil reads the padded buffer without edge cases.
//
// One CTA = 128x64 output tile in PADDED coordinates (so every global store is an aligned 32-bit word; the few frame
// bytes a word may cover are never read by anyone).
//   * TMA stages the 160x70 input box into shared memory (3-px halo; the box starts 16 px left of the tile because a
//     TMA box must start on a 16-byte boundary of the innermost dimension -- an unaligned start faults with 'illegal
//     instruction' on B200, see tools/probe/tma_probe2.cu).
//   * horizontal pass on the integer dot-product unit: out(x) = dp4a(bytes x-3..x, {18,34,48,56}) + dp4a(bytes x+1..x+4,
//     {48,34,18,0}); the eight byte windows of four adjacent outputs come from three aligned words by funnel shifts.  A thread
//     computes the same four columns of TWO consecutive rows and stores the sums (<= 255*256, 16 bits) as u16x2
//     {row 2j, row 2j+1} -- one 128-bit shared store.
//   * vertical pass, also on the dot-product unit: with rows paired like that, the 7 taps of an output row are four
//     dp2a (two 16-bit sums x two 8-bit taps each) on top of the rounding constant; a thread owns 4 columns x 8 rows,
//     reads its 7 row pairs with 128-bit loads and writes one aligned word per row (a warp writes 128 contiguous bytes).
//     ~11 instructions per pixel instead of ~25 for the u16x2 SIMD / 32-bit IMAD version it replaces.
#include "fbe_internal.cuh"
#include "tma.cuh"

namespace fbe {

constexpr int kBlurRawW = kBlurTW + 32, kBlurRawH = kBlurTH + 6;     // 160 x 70 staged bytes
constexpr int kBlurRawWords = kBlurRawW / 4;
constexpr int kBlurPairs = kBlurRawH / 2;                             // 35 row pairs
static_assert(kBlurRawH % 2 == 0 && kBlurTH % 8 == 0, "rows are processed in pairs / bands of 8");

constexpr unsigned kTapsLo = 0x38302212u;        // bytes {18, 34, 48, 56}
constexpr unsigned kTapsHi = 0x00122230u;        // bytes {48, 34, 18, 0}
constexpr unsigned kTapsOddLo = 0x30221200u;     // bytes {0, 18, 34, 48}
constexpr unsigned kTapsOddHi = 0x12223038u;     // bytes {56, 48, 34, 18}

// 7-tap row sums of the four outputs whose first tap is byte 1 of w0 (w0, w1, w2 = 12 consecutive bytes)
__device__ __forceinline__ void blur_hquad(unsigned w0, unsigned w1, unsigned w2, unsigned h[4]) {
    h[0] = __dp4a(__funnelshift_r(w0, w1, 8), kTapsLo, __dp4a(__funnelshift_r(w1, w2, 8), kTapsHi, 0u));
    h[1] = __dp4a(__funnelshift_r(w0, w1, 16), kTapsLo, __dp4a(__funnelshift_r(w1, w2, 16), kTapsHi, 0u));
    h[2] = __dp4a(__funnelshift_r(w0, w1, 24), kTapsLo, __dp4a(__funnelshift_r(w1, w2, 24), kTapsHi, 0u));
    h[3] = __dp4a(w1, kTapsLo, __dp4a(w2, kTapsHi, 0u));
}

__global__ void __launch_bounds__(256) k_blur(const Plan* __restrict__ plan, Workspace ws, const __grid_constant__ TmaMaps maps) {
    __shared__ __align__(128) uint32_t raw[kBlurRawH * kBlurRawWords];
    __shared__ __align__(16) uint32_t hs[kBlurPairs * kBlurTW];           // u16x2: row sums of rows 2j (low) and 2j+1 (high)
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const uint32_t te = __ldg(ws.blur_tab + blockIdx.x);
    const int l = (int)(te >> 24), ty = (int)((te >> 12) & 0xFFFu), tx = (int)(te & 0xFFFu);
    const LevelGeom& g = plan->lv[l];
    const int px0 = tx * kBlurTW;                    // padded column of the tile's first output
    const int y0 = ty * kBlurTH;                     // level row of the tile's first output

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, kBlurRawH * kBlurRawW);
        tma_load_3d(raw, &maps.m[l], &bar, px0 - 16, y0 + kEdge - 3, ws.slot0 + b);
    }
    mbar_wait(&bar, 0);

    // ---- horizontal pass: task = (row pair, quad of 4 output columns) ------------------------------------------------
    for (int k = tid; k < kBlurPairs * (kBlurTW / 4); k += 256) {
        const int pr = k >> 5, q = k & 31;
        // w[0] = raw columns 12 + 4q .. = outputs 4q-4 .. 4q-1 of the tile: the taps of output 4q start at byte 1 of w[0]
        const uint32_t* w = raw + (2 * pr) * kBlurRawWords + q + 3;
        unsigned h0[4], h1[4];
        blur_hquad(w[0], w[1], w[2], h0);
        blur_hquad(w[kBlurRawWords], w[kBlurRawWords + 1], w[kBlurRawWords + 2], h1);
        uint4 o;
        o.x = __byte_perm(h0[0], h1[0], 0x5410); o.y = __byte_perm(h0[1], h1[1], 0x5410);
        o.z = __byte_perm(h0[2], h1[2], 0x5410); o.w = __byte_perm(h0[3], h1[3], 0x5410);
        *reinterpret_cast<uint4*>(hs + pr * kBlurTW + 4 * q) = o;
    }
    __syncthreads();

    // ---- vertical pass: 4 columns x 8 rows per thread ----------------------------------------------------------------
    constexpr int kRows = kBlurTH / 8;               // output rows per thread (8 warps = 8 bands)
    const int cg = tid & 31, band = tid >> 5;
    const int pcol = px0 + 4 * cg;
    if (pcol >= g.pitch) return;
    // output row 8*band + r reads raw rows 8*band + r .. + 6, i.e. row pairs 4*band + (r >> 1) .. + 3
    uint4 pp[kRows / 2 + 3];
#pragma unroll
    for (int j = 0; j < kRows / 2 + 3; ++j) pp[j] = *reinterpret_cast<const uint4*>(hs + (band * (kRows / 2) + j) * kBlurTW + 4 * cg);
    uint8_t* out = ws.blur + (size_t)b * plan->pyr_bytes + g.img_off + pcol;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int y = y0 + band * kRows + r;
        if (y >= g.h) break;
        const int m = r >> 1;
        unsigned a[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned p0 = c == 0 ? pp[m].x : c == 1 ? pp[m].y : c == 2 ? pp[m].z : pp[m].w;
            const unsigned p1 = c == 0 ? pp[m + 1].x : c == 1 ? pp[m + 1].y : c == 2 ? pp[m + 1].z : pp[m + 1].w;
            const unsigned p2 = c == 0 ? pp[m + 2].x : c == 1 ? pp[m + 2].y : c == 2 ? pp[m + 2].z : pp[m + 2].w;
            const unsigned p3 = c == 0 ? pp[m + 3].x : c == 1 ? pp[m + 3].y : c == 2 ? pp[m + 3].z : pp[m + 3].w;
            if ((r & 1) == 0)        // rows 2m .. 2m+6: taps {18,34} {48,56} {48,34} {18,0}
                a[c] = __dp2a_hi(p3, kTapsHi, __dp2a_lo(p2, kTapsHi, __dp2a_hi(p1, kTapsLo, __dp2a_lo(p0, kTapsLo, 32768u))));
            else                     // rows 2m+1 .. 2m+7: taps {0,18} {34,48} {56,48} {34,18}
                a[c] = __dp2a_hi(p3, kTapsOddHi, __dp2a_lo(p2, kTapsOddHi, __dp2a_hi(p1, kTapsOddLo, __dp2a_lo(p0, kTapsOddLo, 32768u))));
        }
        // byte 2 of every sum is the result (sum < 2^24)
        const unsigned lo = __byte_perm(a[0], a[1], 0x0062), hi = __byte_perm(a[2], a[3], 0x0062);
        *reinterpret_cast<uint32_t*>(out + (size_t)(y + kEdge) * g.pitch) = __byte_perm(lo, hi, 0x5410);
    }
}

int launch_blur(const Plan& hp, const Plan* dp, const Workspace& ws, const TmaMaps& maps, int nimg, cudaStream_t st) {
    dim3 grid(hp.blur_tiles_total, nimg);
    k_blur<<<grid, 256, 0, st>>>(dp, ws, maps);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
