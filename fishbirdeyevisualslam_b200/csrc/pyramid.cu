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

// One CTA = 128 output columns x 64 output rows of one padded level; a thread owns 4 adjacent columns and walks 8 rows.
// The column coefficients (source offset + two 11-bit weights) stay in registers for the whole walk; the horizontal
// interpolation of a source row is kept and reused when the next output row needs it again (scale 1.2: consecutive
// output rows share one of their two source rows most of the time).  Source bytes come through L1/L2 (each source
// pixel is read by ~1.4 threads); stores are aligned 32-bit words, 128 contiguous bytes per warp.
constexpr int kRsRowsPerWarp = 8, kRsWarps = 8;

__global__ void __launch_bounds__(256) k_resize(const Plan* __restrict__ plan, Workspace ws,
                                                const ResizeTab* __restrict__ tab, int level) {
    const LevelGeom& g = plan->lv[level];
    const LevelGeom& s = plan->lv[level - 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x4 = (blockIdx.x * 32 + lane) * 4;
    const int b = blockIdx.z;
    const int pitch = g.pitch, spitch = s.pitch, ph = g.ph;
    if (x4 >= pitch) return;
    const uint8_t* src = ws.pyr + (size_t)b * plan->pyr_bytes + s.img_off;
    uint8_t* dst = ws.pyr + (size_t)b * plan->pyr_bytes + g.img_off + x4;
    // 4 column entries = 32 bytes, 32-byte aligned
    const uint4* tx4 = reinterpret_cast<const uint4*>(tab + g.tabx_off + x4);
    const uint4 ta = __ldg(tx4), tb = __ldg(tx4 + 1);
    const int ofs[4] = {(int)ta.x, (int)ta.z, (int)tb.x, (int)tb.z};
    const int a0[4] = {(int)(short)(ta.y & 0xFFFFu), (int)(short)(ta.w & 0xFFFFu), (int)(short)(tb.y & 0xFFFFu), (int)(short)(tb.w & 0xFFFFu)};
    const int a1[4] = {(int)ta.y >> 16, (int)ta.w >> 16, (int)tb.y >> 16, (int)tb.w >> 16};
    const ResizeTab* ty = tab + g.taby_off;
    int y = blockIdx.y * (kRsRowsPerWarp * kRsWarps) + wid * kRsRowsPerWarp;
    const int yend = min(y + kRsRowsPerWarp, ph);
    int cur = -4;                 // source row whose interpolation sits in r_lo (r_hi holds cur + 1)
    int r_lo[4], r_hi[4];
    auto hrow = [&](int sy, int (&r)[4]) {
        const uint8_t* S = src + (size_t)sy * spitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = ((int)S[ofs[i]] * a0[i] + (int)S[ofs[i] + 1] * a1[i]) >> 4;
    };
    for (; y < yend; ++y) {
        const ResizeTab t = ty[y];
        const int sy = t.ofs;
        if (sy == cur + 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) r_lo[i] = r_hi[i];
            hrow(sy + 1, r_hi);      // row sy+1 exists in the padded source (frame), weight 0 when clamped
        } else if (sy != cur) {
            hrow(sy, r_lo);
            hrow(sy + 1, r_hi);
        }
        cur = sy;
        const int b0 = t.a0, b1 = t.a1;
        uint32_t out = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int v = (((b0 * r_lo[i]) >> 16) + ((b1 * r_hi[i]) >> 16) + 2) >> 2;
            v = min(max(v, 0), 255);
            out |= (uint32_t)v << (8 * i);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)y * pitch) = out;
    }
}

int launch_pyramid(const Plan& hp, const Plan* dp, const Workspace& ws, const ResizeTab* d_tab, int nimg, cudaStream_t st) {
    for (int l = 0; l < hp.nlevels; ++l) {
        const LevelGeom& g = hp.lv[l];
        dim3 block(64);
        if (g.pitch / 4 > 64) block.x = 128;
        if (g.pitch / 4 > 128) block.x = 256;
        dim3 grid((g.pitch / 4 + block.x - 1) / block.x, g.ph, nimg);
        if (l == 0) k_level0<<<grid, block, 0, st>>>(dp, ws);
        else {
            dim3 rgrid((g.pitch / 4 + 31) / 32, (g.ph + kRsRowsPerWarp * kRsWarps - 1) / (kRsRowsPerWarp * kRsWarps), nimg);
            k_resize<<<rgrid, 256, 0, st>>>(dp, ws, d_tab, l);
        }
        count_launch();
    }
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
