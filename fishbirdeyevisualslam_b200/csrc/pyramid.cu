// ComputePyramid (src/ORBextractor.cc:1107-1132) as two sm_100a kernels.
//   k_level0 : copyMakeBorder(image, REFLECT_101) into the padded level-0 buffer
//   k_resize : level l from level l-1 -- cv::resize INTER_LINEAR 8U (11-bit fixed point, separable) evaluated for
//              every PADDED destination pixel: frame pixels are the bilinear result at their REFLECT_101 source, so
//              resize + copyMakeBorder collapse into one pass and one aligned 4-byte store per thread.
// HBM-bound integer work: each thread produces 4 horizontally adjacent bytes; rows are 16-byte aligned.
#include "fbe_internal.cuh"

namespace fbe {

__device__ __forceinline__ int reflect101_dev(int p, int len) {
    // |p| excursion is at most 19 and len >= 20 for every level that passes build_plan, one fold suffices;
    // the loop keeps it correct for any len.
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256) k_level0(const Plan* __restrict__ plan, Workspace ws) {
    const LevelGeom g = plan->lv[0];
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    const int b = blockIdx.z;
    if (x4 >= g.pitch) return;
    const uint8_t* src = ws.in + (size_t)b * ws.in_slot_stride;
    const int ys = reflect101_dev(y - kEdge, g.h);
    const uint8_t* row = src + (size_t)ys * ws.in_pitch;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int x = x4 + i;
        uint32_t px = 0;
        if (x < g.w + 2 * kEdge) px = row[reflect101_dev(x - kEdge, g.w)];
        v |= px << (8 * i);
    }
    uint8_t* dst = ws.pyr + (size_t)b * plan->pyr_bytes + g.img_off;
    *reinterpret_cast<uint32_t*>(dst + (size_t)y * g.pitch + x4) = v;
}

__global__ void __launch_bounds__(256) k_resize(const Plan* __restrict__ plan, Workspace ws,
                                                const ResizeTab* __restrict__ tab, int level) {
    const LevelGeom g = plan->lv[level];
    const LevelGeom s = plan->lv[level - 1];
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    const int b = blockIdx.z;
    if (x4 >= g.pitch) return;
    const uint8_t* src = ws.pyr + (size_t)b * plan->pyr_bytes + s.img_off;
    const ResizeTab ty = tab[g.taby_off + y];
    const uint8_t* S0 = src + (size_t)ty.ofs * s.pitch;
    const uint8_t* S1 = S0 + s.pitch;       // row sy+1 exists in the padded source (frame), weight 0 when clamped
    const int b0 = ty.a0, b1 = ty.a1;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const ResizeTab tx = tab[g.tabx_off + x4 + i];
        const int a0 = tx.a0, a1 = tx.a1;
        const int r0 = S0[tx.ofs] * a0 + S0[tx.ofs + 1] * a1;
        const int r1 = S1[tx.ofs] * a0 + S1[tx.ofs + 1] * a1;
        int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        out |= (uint32_t)v << (8 * i);
    }
    uint8_t* dst = ws.pyr + (size_t)b * plan->pyr_bytes + g.img_off;
    *reinterpret_cast<uint32_t*>(dst + (size_t)y * g.pitch + x4) = out;
}

int launch_pyramid(const Plan& hp, const Plan* dp, const Workspace& ws, const ResizeTab* d_tab, int nimg, cudaStream_t st) {
    for (int l = 0; l < hp.nlevels; ++l) {
        const LevelGeom& g = hp.lv[l];
        dim3 block(64);
        if (g.pitch / 4 > 64) block.x = 128;
        if (g.pitch / 4 > 128) block.x = 256;
        dim3 grid((g.pitch / 4 + block.x - 1) / block.x, g.ph, nimg);
        if (l == 0) k_level0<<<grid, block, 0, st>>>(dp, ws);
        else k_resize<<<grid, block, 0, st>>>(dp, ws, d_tab, l);
        count_launch();
    }
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
