// ComputePyramid (src/ORBextractor.cc:1107-1132) as two sm_100a kernels.
//   k_level0 : copyMakeBorder(image, REFLECT_101) into the padded level-0 buffer
//   k_resize : level l from level l-1 -- cv::resize INTER_LINEAR 8U (11-bit fixed point, separable) evaluated for
//              every PADDED destination pixel: frame pixels are the bilinear result at their REFLECT_101 source, so
//              resize + copyMakeBorder collapse into one pass and one aligned 4-byte store per thread.
// HBM-bound integer work: each thread produces 4 horizontally adjacent bytes; rows are 16-byte aligned.
#include "fbe_internal.cuh"
#include "tma.cuh"

namespace fbe {


// Level 0 in two launches so that no warp mixes the two paths:
//   k_level0       : interior 16-byte chunks (destination column d holds source column d - 19, i.e. the source is displaced
//                    by 3 bytes modulo 4): one 32-bit + one 128-bit aligned load of the source row, realigned by a funnel
//                    shift, one 128-bit store.  Needs a 16-byte aligned source (base, pitch, image stride).
//   k_level0_border: the chunks that touch the reflected frame (and every chunk of an unaligned source): per-byte
//                    REFLECT_101 gather (one fold suffices: the excursion is 19 and every level is at least 20 px).
__device__ __forceinline__ int reflect101_once(int p, int len) {
    p = p < 0 ? -p : p;
    return p >= len ? 2 * (len - 1) - p : p;
}

__global__ void __launch_bounds__(256) k_level0(const Plan* __restrict__ plan, Workspace ws, int first_chunk, int nchunks) {
    pdl_launch_dependents();
    pdl_wait();
    const LevelGeom& g = plan->lv[0];
    const int ci = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (ci >= nchunks || y >= g.ph) return;
    const int c16 = (first_chunk + ci) * 16;
    const uint8_t* row = ws.in + (size_t)b * ws.in_slot_stride + (size_t)reflect101_once(y - kEdge, g.h) * ws.in_pitch;
    const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(row + c16 - 20));
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + c16 - 16));
    uint4 o;
    o.x = __funnelshift_r(w0, q.x, 8); o.y = __funnelshift_r(q.x, q.y, 8);
    o.z = __funnelshift_r(q.y, q.z, 8); o.w = __funnelshift_r(q.z, q.w, 8);
    uint8_t* dst = ws.pyr + (size_t)b * plan->pyr_bytes + g.img_off;
    *reinterpret_cast<uint4*>(dst + (size_t)y * g.pitch + c16) = o;
}

// border chunks: chunk index = bc < nlead ? bc : first_tail + (bc - nlead)
__global__ void __launch_bounds__(256) k_level0_border(const Plan* __restrict__ plan, Workspace ws, int nlead, int first_tail, int nborder) {
    pdl_launch_dependents();
    pdl_wait();
    const LevelGeom& g = plan->lv[0];
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.z;
    const int y = t / nborder, bc = t - y * nborder;
    if (y >= g.ph) return;
    const int c16 = (bc < nlead ? bc : first_tail + (bc - nlead)) * 16;
    const int w = g.w;
    const uint8_t* row = ws.in + (size_t)b * ws.in_slot_stride + (size_t)reflect101_once(y - kEdge, g.h) * ws.in_pitch;
    uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int x = c16 + i;
        uint32_t px = 0;
        if (x < w + 2 * kEdge) px = row[reflect101_once(x - kEdge, w)];
        v[i >> 2] |= px << (8 * (i & 3));
    }
    uint8_t* dst = ws.pyr + (size_t)b * plan->pyr_bytes + g.img_off;
    *reinterpret_cast<uint4*>(dst + (size_t)y * g.pitch + c16) = make_uint4(v[0], v[1], v[2], v[3]);
}

// One CTA = 128 output columns x 64 output rows of one padded level; a thread owns 4 adjacent columns and walks 8 rows.
// The column coefficients (source offset + two 11-bit weights) stay in registers for the whole walk; the horizontal
// interpolation of a source row is kept and reused when the next output row needs it again (scale 1.2: consecutive
// output rows share one of their two source rows most of the time).  Source bytes come through L1/L2 (each source
// pixel is read by ~1.4 threads); stores are aligned 32-bit words, 128 contiguous bytes per warp.
constexpr int kRsRowsPerWarp = 8, kRsWarps = 8;

__global__ void __launch_bounds__(256) k_resize(const Plan* __restrict__ plan, Workspace ws,
                                                const ResizeTab* __restrict__ tab, int level) {
    pdl_launch_dependents();
    pdl_wait();
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

// TMA-staged variant (used whenever the source tile of a 128x64 output tile fits one TMA box, i.e. for every scale
// factor up to ~1.8): the source tile lands in shared memory, so a source byte costs one LDS with a 32-bit address
// instead of a 64-bit global address computation; (b*r) >> 16 is IMAD + SHF (IMAD.HI measured at 0.4x the IMAD rate on B200), and the
// saturation of the reference formula is provably never reached (weights sum to 2047..2049), so the
// combine is 2 IMAD + 2 SHF + IADD3 + SHF per pixel.
__global__ void __launch_bounds__(256) k_resize_tma(const Plan* __restrict__ plan, Workspace ws, const ResizeTab* __restrict__ tab,
                                                    int level, const __grid_constant__ CUtensorMap map) {
    extern __shared__ __align__(128) uint8_t rs_tile[];
    __shared__ __align__(8) uint64_t bar;
    const LevelGeom& g = plan->lv[level];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int bw = g.rs_bw, bh = g.rs_bh;
    const int x0s = plan->rs_x0[level][blockIdx.x], y0s = plan->rs_y0[level][blockIdx.y];
    pdl_launch_dependents();
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        pdl_wait();              // the source level: everything the CTA reads of it arrives through this TMA load; all threads wait on its barrier
        mbar_expect_tx(&bar, (uint32_t)(bw * bh));
        tma_load_3d(rs_tile, &map, &bar, x0s, y0s, ws.slot0 + b);
    }
    const int x4 = (blockIdx.x * 32 + lane) * 4;
    const int pitch = g.pitch, ph = g.ph;
    const bool active = x4 < pitch;
    // The four columns of a thread read source bytes within a span of <= 8 (scale factors up to 1.8, reflected frame columns
    // included): three aligned words of the staged row cover it.  base4 / sh locate the span, sel[i] picks the byte pair
    // (S[col], S[col + 1]) of column i out of the realigned 8 bytes, w01[i] holds its two 11-bit weights as u16x2 for IDP.2A.
    // (build_plan sends a level to the direct-global kernel when a group of four columns spans more than that.)
    int base4 = 0, sh = 0;
    unsigned sel[4] = {0, 0, 0, 0}, w01[4] = {0, 0, 0, 0};
    if (active) {
        int col[4], a0[4], a1[4];
        const uint4* tx4 = reinterpret_cast<const uint4*>(tab + g.tabx_off + x4);      // 4 column entries = 32 bytes
        const uint4 ta = __ldg(tx4), tb = __ldg(tx4 + 1);
        col[0] = (int)ta.x - x0s; col[1] = (int)ta.z - x0s; col[2] = (int)tb.x - x0s; col[3] = (int)tb.z - x0s;
        a0[0] = (int)(short)(ta.y & 0xFFFFu); a0[1] = (int)(short)(ta.w & 0xFFFFu); a0[2] = (int)(short)(tb.y & 0xFFFFu); a0[3] = (int)(short)(tb.w & 0xFFFFu);
        a1[0] = (int)ta.y >> 16; a1[1] = (int)ta.w >> 16; a1[2] = (int)tb.y >> 16; a1[3] = (int)tb.w >> 16;
        const int cmin = min(min(col[0], col[1]), min(col[2], col[3]));
        base4 = cmin & ~3; sh = (cmin & 3) * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int d = col[i] - cmin;
            sel[i] = (unsigned)d | ((unsigned)(d + 1) << 4);
            w01[i] = (unsigned)a0[i] | ((unsigned)a1[i] << 16);               // weights are in [0, 2048]
        }
    }
    mbar_wait(&bar, 0);
    if (!active) return;
    uint8_t* dst = ws.pyr + (size_t)b * plan->pyr_bytes + g.img_off + x4;
    const ResizeTab* ty = tab + g.taby_off;
    int y = blockIdx.y * kRsTH + wid * kRsRowsPerWarp;
    const int yend = min(y + kRsRowsPerWarp, ph);
    int cur = -4;                 // source row (tile coordinates) whose interpolation sits in r_lo (r_hi holds cur + 1)
    unsigned r_lo[4], r_hi[4];
    auto hrow = [&](int sy, unsigned (&r)[4]) {
        // 3 LDS.32 + 2 SHF realign the span; per column one PRMT (byte pair) + one IDP.2A (a0 * S0 + a1 * S1) + the >> 4
        const uint32_t* W = reinterpret_cast<const uint32_t*>(rs_tile + sy * bw + base4);
        const uint32_t w0 = W[0], w1 = W[1], w2 = W[2];
        const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = __dp2a_lo(w01[i], __byte_perm(lo, hi, sel[i]), 0u) >> 4;
    };
    for (; y < yend; ++y) {
        const ResizeTab t = ty[y];
        const int sy = t.ofs - y0s;
        if (sy == cur + 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) r_lo[i] = r_hi[i];
            hrow(sy + 1, r_hi);      // row sy+1 exists in the padded source (frame), weight 0 when clamped
        } else if (sy != cur) {
            hrow(sy, r_lo);
            hrow(sy + 1, r_hi);
        }
        cur = sy;
        const unsigned b0 = (unsigned)(int)t.a0, b1 = (unsigned)(int)t.a1;                   // weights are in [0, 2048], r < 2^15: products fit 32 bits
        unsigned v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (((b0 * r_lo[i]) >> 16) + ((b1 * r_hi[i]) >> 16) + 2u) >> 2;   // <= 255, see above; IMAD + SHF: IMAD.HI runs at 0.4x the IMAD rate
        *reinterpret_cast<uint32_t*>(dst + (size_t)y * pitch) = v[0] + (v[1] << 8) + (v[2] << 16) + (v[3] << 24);
    }
}

int launch_pyramid(const Plan& hp, const Plan* dp, const Workspace& ws, const ResizeTab* d_tab, const TmaMaps& rs_maps, int nimg,
                   cudaStream_t st) {
    for (int l = 0; l < hp.nlevels; ++l) {
        const LevelGeom& g = hp.lv[l];
        if (l == 0) {
            const int aligned16 = ((reinterpret_cast<uintptr_t>(ws.in) | (uintptr_t)ws.in_pitch | (uintptr_t)ws.in_slot_stride) & 15) == 0;
            const int nchunk = g.pitch / 16;
            // interior chunks: source bytes c16-20 .. c16-1 all inside the row  <=>  32 <= c16 <= w
            const int first = 2, last = std::min(g.w / 16, nchunk - 1);
            const int nint = (aligned16 && last >= first) ? last - first + 1 : 0;
            if (nint > 0) {
                dim3 grid((nint + 31) / 32, (g.ph + 7) / 8, nimg);
                FBE_CUDA(launch_dep(k_level0, grid, dim3(32, 8), 0, st, dp, ws, first, nint));
                count_launch();
            }
            const int nlead = nint > 0 ? first : nchunk, first_tail = last + 1;
            const int nborder = nint > 0 ? nlead + (nchunk - first_tail) : nchunk;
            dim3 bgrid((g.ph * nborder + 255) / 256, 1, nimg);
            FBE_CUDA(launch_dep(k_level0_border, bgrid, dim3(256), 0, st, dp, ws, nlead, first_tail, nborder));
        } else {
            dim3 rgrid((g.pitch + kRsTW - 1) / kRsTW, (g.ph + kRsTH - 1) / kRsTH, nimg);
            if (g.rs_bw > 0) {
                const size_t smem = (size_t)g.rs_bw * g.rs_bh + 16;     // + 16: the word loads of the last row's last span may run past the box
                if (smem > 48 * 1024) FBE_CUDA(cudaFuncSetAttribute(k_resize_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 16));   // box <= 64 KB (build_plan) + the slack
                FBE_CUDA(launch_dep(k_resize_tma, rgrid, dim3(256), smem, st, dp, ws, d_tab, l, rs_maps.m[l]));
            } else {
                FBE_CUDA(launch_dep(k_resize, rgrid, dim3(256), 0, st, dp, ws, d_tab, l));
            }
        }
        count_launch();
    }
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
