// GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of every pyramid level (src/ORBextractor.cc:1085-1086).
// OpenCV >= 3.4.2 evaluates this 8-bit case exactly: integer kernel [18,34,48,56,48,34,18]/256 per axis, no
// intermediate rounding, out = (sum_y sum_x k_y k_x p + 32768) >> 16 (pinned against cv2 4.13).
// The reference blurs a clone of the un-padded level with REFLECT_101 at its edge; our padded level already
// carries that reflection in its 19-px frame, so the stencil reads the padded buffer without edge cases.
// One CTA = 64x16 output tile; the 70x22 input tile is staged in shared memory, the horizontal pass leaves a
// 64x22 u16 intermediate in shared memory, the vertical pass writes 4 bytes per thread.
#include "fbe_internal.cuh"

namespace fbe {

constexpr int kBlurTW = 64, kBlurTH = 16;

__global__ void __launch_bounds__(256) k_blur(const Plan* __restrict__ plan, Workspace ws, int level) {
    __shared__ __align__(16) uint8_t tile[(kBlurTH + 6) * 72];
    __shared__ __align__(16) uint16_t hsum[(kBlurTH + 6) * kBlurTW];
    const LevelGeom g = plan->lv[level];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kBlurTW, y0 = blockIdx.y * kBlurTH;     // level coordinates of the tile
    const uint8_t* img = ws.pyr + (size_t)b * plan->pyr_bytes + g.img_off;
    uint8_t* out = ws.blur + (size_t)b * plan->pyr_bytes + g.img_off;
    // stage rows y0-3 .. y0+TH+2, cols x0-3 .. x0+TW+2 (padded coords: +19); clamp reads to the padded extent
    const int pw = g.w + 2 * kEdge;
    for (int i = threadIdx.x; i < (kBlurTH + 6) * 70; i += 256) {
        const int ty = i / 70, tx = i - ty * 70;
        const int py = min(y0 - 3 + ty + kEdge, g.ph - 1), px = min(x0 - 3 + tx + kEdge, pw - 1);
        tile[ty * 72 + tx] = img[(size_t)py * g.pitch + px];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (kBlurTH + 6) * kBlurTW; i += 256) {
        const int ty = i / kBlurTW, tx = i - ty * kBlurTW;
        const uint8_t* p = tile + ty * 72 + tx;
        hsum[i] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    // 64x16 outputs, 4 per thread (one aligned-in-tile 32-bit store)
    const int tx4 = (threadIdx.x & 15) * 4, ty = threadIdx.x >> 4;
    const int y = y0 + ty;
    if (y >= g.h) return;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint16_t* q = hsum + ty * kBlurTW + tx4 + i;
        const uint32_t a = 18u * (q[0] + q[6 * kBlurTW]) + 34u * (q[kBlurTW] + q[5 * kBlurTW]) +
                           48u * (q[2 * kBlurTW] + q[4 * kBlurTW]) + 56u * q[3 * kBlurTW];
        v |= ((a + 32768u) >> 16) << (8 * i);
    }
    uint8_t* dst = out + (size_t)(y + kEdge) * g.pitch + (x0 + tx4 + kEdge);
    const int x = x0 + tx4;
    if (x + 3 < g.w) {
        dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); dst[3] = (uint8_t)(v >> 24);
    } else {
        for (int i = 0; i < 4 && x + i < g.w; ++i) dst[i] = (uint8_t)(v >> (8 * i));
    }
}

int launch_blur(const Plan& hp, const Plan* dp, const Workspace& ws, int nimg, cudaStream_t st) {
    for (int l = 0; l < hp.nlevels; ++l) {
        const LevelGeom& g = hp.lv[l];
        dim3 grid((g.w + kBlurTW - 1) / kBlurTW, (g.h + kBlurTH - 1) / kBlurTH, nimg);
        k_blur<<<grid, 256, 0, st>>>(dp, ws, l);
        count_launch();
    }
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
