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
//   * horizontal pass, u16x2 SIMD: a row sum is <= 255*256 < 2^16, so two pixels share one 32-bit IMAD; the byte pairs
//     come from PRMT/SHF on three aligned words.  4 outputs per thread task -> one 64-bit shared store.
//   * vertical pass, 32-bit: a thread owns 4 columns x 8 rows, keeps its 14 input rows in registers, and writes one
//     aligned word per row (a warp writes 128 contiguous bytes).
#include "fbe_internal.cuh"
#include "tma.cuh"

namespace fbe {

constexpr int kBlurRawW = kBlurTW + 32, kBlurRawH = kBlurTH + 6;     // 160 x 70 staged bytes
constexpr int kBlurRawWords = kBlurRawW / 4;

__device__ __forceinline__ unsigned blur_hpair(unsigned a, unsigned b, unsigned c, unsigned d, unsigned e, unsigned f, unsigned g) {
    return 18u * (a + g) + 34u * (b + f) + 48u * (c + e) + 56u * d;
}

__global__ void __launch_bounds__(256) k_blur(const Plan* __restrict__ plan, Workspace ws, const __grid_constant__ TmaMaps maps) {
    __shared__ __align__(128) uint32_t raw[kBlurRawH * kBlurRawWords];
    __shared__ __align__(16) uint32_t hs[kBlurRawH * (kBlurTW / 2)];      // u16x2: row sums of two adjacent pixels
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

    // ---- horizontal pass: task = (row, quad of 4 output columns) ---------------------------------------------------
    for (int k = tid; k < kBlurRawH * (kBlurTW / 4); k += 256) {
        const int row = k >> 5, q = k & 31;
        const uint32_t* w = raw + row * kBlurRawWords + q + 3;  // w[0] = columns 4q-4 .. 4q-1 relative to the tile's outputs
        const unsigned w0 = w[0], w1 = w[1], w2 = w[2];
        // p0..p11 = bytes of w0,w1,w2; outputs 0,1 use taps p1..p7 / p2..p8, outputs 2,3 use p3..p9 / p4..p10
        const unsigned s1 = __funnelshift_r(w0, w1, 8), s2 = __funnelshift_r(w1, w2, 8);   // p1..p4, p5..p8
        const unsigned O0 = __byte_perm(s1, 0, 0x4140), O1 = __byte_perm(s1, 0, 0x4342);   // (p1,p2) (p3,p4)
        const unsigned O2 = __byte_perm(s2, 0, 0x4140), O3 = __byte_perm(s2, 0, 0x4342);   // (p5,p6) (p7,p8)
        const unsigned O4 = __byte_perm(w2, 0, 0x4241);                                   // (p9,p10)
        const unsigned E1 = __byte_perm(w0, 0, 0x4342), E2 = __byte_perm(w1, 0, 0x4140);   // (p2,p3) (p4,p5)
        const unsigned E3 = __byte_perm(w1, 0, 0x4342), E4 = __byte_perm(w2, 0, 0x4140);   // (p6,p7) (p8,p9)
        uint2 o;
        o.x = blur_hpair(O0, E1, O1, E2, O2, E3, O3);
        o.y = blur_hpair(O1, E2, O2, E3, O3, E4, O4);
        *reinterpret_cast<uint2*>(hs + row * (kBlurTW / 2) + 2 * q) = o;
    }
    __syncthreads();

    // ---- vertical pass: 4 columns x 8 rows per thread ----------------------------------------------------------------
    constexpr int kRows = kBlurTH / 8;               // output rows per thread (8 warps = 8 bands)
    const int cg = tid & 31, band = tid >> 5;
    const int pcol = px0 + 4 * cg;
    if (pcol >= g.pitch) return;
    unsigned h[kRows + 6][4];
#pragma unroll
    for (int r = 0; r < kRows + 6; ++r) {
        const uint2 v = *reinterpret_cast<const uint2*>(hs + (band * kRows + r) * (kBlurTW / 2) + 2 * cg);
        h[r][0] = v.x & 0xFFFFu; h[r][1] = v.x >> 16; h[r][2] = v.y & 0xFFFFu; h[r][3] = v.y >> 16;
    }
    uint8_t* out = ws.blur + (size_t)b * plan->pyr_bytes + g.img_off + pcol;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int y = y0 + band * kRows + r;
        if (y >= g.h) break;
        unsigned a[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            a[c] = 18u * (h[r][c] + h[r + 6][c]) + 34u * (h[r + 1][c] + h[r + 5][c]) + 48u * (h[r + 2][c] + h[r + 4][c]) +
                   56u * h[r + 3][c] + 32768u;
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
