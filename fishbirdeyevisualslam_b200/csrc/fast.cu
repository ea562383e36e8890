// Per-cell FAST-9/16 detection with threshold fallback (ComputeKeyPointsOctTree cell loop,
// src/ORBextractor.cc:765-829) as ONE kernel.  A CTA owns a GROUP of up to 8 consecutive ~30x30 FAST cells of one
// cell row of one level of one image (a strip <= 224 px wide), so that the 3-px ring halo and the per-CTA fixed costs
// are amortised over ~8k pixels.
//
// What the reference does per cell: cv::FAST(roi, iniThFAST, nms) and, when that returns nothing,
// cv::FAST(roi, minThFAST, nms).  Facts used (pinned against cv2 by the oracle tests):
//   * FAST score of a pixel = m - 1, m = max over the 16 arcs of 9 contiguous ring pixels of min |I - c| with one
//     sign; corner at threshold t  <=>  m > t.
//   * pixels in the 3-px frame of the ROI are never corners, so a cell's detection region is
//     [19 + j*wCell, 19 + (j+1)*wCell) x [19 + i*hCell, ...) clipped to [19, cols-19) x [19, rows-19): regions tile
//     the level exactly and NMS never looks across a cell edge (outside counts as score 0).
//   * FAST(t) == { k in FAST(lo) : k.response >= t } for t >= lo on the same ROI, so one score/NMS pass at the low
//     threshold yields both answers; the fallback decision is "no NMS survivor with score >= iniThFAST".
//
// Phases of a CTA (all in shared memory, 4 block barriers):
//   0. TMA stages the strip + 3-px halo (box start rounded down to 16 bytes); the score tile and the survivor masks
//      are cleared while the copy is in flight.
//   1. PRETEST every pixel (8 per thread) with the 4 compass ring pixels (every 9-arc contains k or k+8 for each k, so
//      min(max(I0,I8),max(I4,I12)) - c > t  or  c - max(min(I0,I8),min(I4,I12)) > t is necessary); passing pixels are
//      appended to the warp's private work list (ballot ranks, no atomics) -- this removes the divergence of the score phase.
//   2. SCORE the work list densely: ring differences are packed as biased s16x2 {I-c+256, c-I+256} with ONE IMAD each, so
//      that bright and dark arcs share the DPX 3-input min/max (VIMNMX3.S16x2): 16+16 min3 + 8 max3 per pixel.  The corners
//      (m > minThFAST) get their score written to the score tile and are compacted IN PLACE to the front of the list.
//   3. strict 8-neighbour NMS of the corners inside their cell (dense: no divergence on non-corners; cells are separated by
//      a zero column in the score tile, so there are no edge cases); survivors set a bit in per-cell row masks (all
//      survivors / survivors with score >= iniThFAST).
//   4. EMIT: one warp per cell turns the row masks into the (y, x)-ordered slot list of the cell (ballot-free: popc +
//      warp scan), choosing the iniThFAST mask when it is non-empty, else the minThFAST one.
// Output per cell: survivors in (y, x) order as packed keys in the cell's private slot range + a count.  The octree
// kernel concatenates cells in row-major order, which reproduces vToDistributeKeys order.
#include <algorithm>
#include "fbe_internal.cuh"
#include "tma.cuh"

namespace fbe {

constexpr int kFastThreads = 256;
constexpr int kFastWarps = kFastThreads / 32;

constexpr int kTilePitch = 256;                             // TMA box width: staged tile column t = padded column tcol0 + t
constexpr int kFastMaxCells = 7;                            // cells are >= 30 px wide, a strip <= 224 px
constexpr int kScorePitch = 256;                            // score tile: strip pixel x of cell cj lives in column x + cj + 1 (one zero column
                                                            // per cell edge, so NMS needs no edge cases); pitch == tile pitch, so a work-list
                                                            // entry (py << 8 | px) addresses both tiles
constexpr int kTileW = kTilePitch / 4;                      // in 32-bit words
static_assert(kFastGroupW + 6 + 15 + 3 <= kTilePitch, "strip + ring halo + 16-byte alignment slack must fit the TMA box");
static_assert(kFastGroupW + kFastMaxCells + 2 <= kScorePitch, "strip + one zero column per cell edge must fit the score tile");

struct FastLayout { int off_sc, sc_bytes, off_work, work_cap, off_mask, mask_words, off_xinfo, total; };

// shared-memory carve-up for a strip of gw x ch pixels covering ncell cells in a level whose cells are hcell high
// (pitches are compile-time constants so that every ring / neighbour access is a base register + immediate)
__host__ __device__ inline FastLayout fast_layout(int gw, int ch, int hcell, int ncell) {
    FastLayout L;
    int o = kTilePitch * (hcell + 6);                // the TMA box always has the level's full cell height
    L.off_sc = o;
    L.sc_bytes = (kScorePitch * (ch + 2) + 15) & ~15;
    o += L.sc_bytes;
    L.off_work = o;
    L.work_cap = ((ch + kFastWarps - 1) / kFastWarps) * gw;          // per warp: its rows, every pixel
    o += (L.work_cap * kFastWarps * 2 + 15) & ~15;
    L.off_mask = o;
    L.mask_words = ncell * ch * 4;                   // [lo|ini][cell][row][2] u32
    o += L.mask_words * 4;
    L.off_xinfo = o;
    o += (gw + 15) & ~15;
    L.total = o;
    return L;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// compass pretest of two pixels held as u16x2 lanes; returns bit 15 / bit 31 set for the lanes that pass
__device__ __forceinline__ unsigned pretest_pair(unsigned c, unsigned i0, unsigned i8, unsigned i4, unsigned i12, unsigned k2) {
    const unsigned e = __vminu2(__vmaxu2(i0, i8), __vmaxu2(i4, i12));     // bright: e - c > t
    const unsigned f = __vmaxu2(__vminu2(i0, i8), __vminu2(i4, i12));     // dark:   c - f > t
    const unsigned d1 = e + 0x01000100u - c;                               // halves in [1, 511]: no borrow between lanes
    const unsigned d2 = c + 0x01000100u - f;
    return (__vmaxu2(d1, d2) + k2) & 0x80008000u;                          // lane > 256 + t  <=>  bit 15 of lane + k set
}

__global__ void __launch_bounds__(kFastThreads, 4) k_fast_cells(const Plan* __restrict__ plan, Workspace ws,
                                                             const __grid_constant__ TmaMaps maps) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;

    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.y;
    const uint32_t te = __ldg(ws.fast_tab + blockIdx.x);
    const int l = (int)(te >> 24), ci = (int)((te >> 12) & 0xFFFu), gj = (int)(te & 0xFFFu);
    const LevelGeom& g = plan->lv[l];
    const int wcell = g.wcell, hcell = g.hcell, ncols = g.ncols;
    const int cj0 = gj * g.gcells, ncell = min(g.gcells, ncols - cj0);

    // detection region of the strip in level coordinates
    const int x0 = kEdge + cj0 * wcell, y0 = kEdge + ci * hcell;
    const int x1 = min(x0 + ncell * wcell, g.w - kEdge), y1 = min(y0 + hcell, g.h - kEdge);
    const int gw = x1 - x0, ch = y1 - y0;
    const int cell0 = g.cell_base + ci * ncols + cj0;
    int* count_out = ws.cell_count + (size_t)b * plan->ncells_total + cell0;
    if (gw <= 0 || ch <= 0) {
        if (tid < ncell) count_out[tid] = 0;
        return;
    }
    const FastLayout L = fast_layout(gw, ch, hcell, ncell);
    uint8_t* tile = smem;
    const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(smem);
    uint8_t* sc = smem + L.off_sc;
    uint16_t* work = reinterpret_cast<uint16_t*>(smem + L.off_work) + wid * L.work_cap;   // this warp's private list
    uint32_t* mask = reinterpret_cast<uint32_t*>(smem + L.off_mask);      // lo masks, then ini masks
    uint8_t* xinfo = smem + L.off_xinfo;

    // ---- phase 0: TMA stages the strip + 3-px ring halo; the box starts on a 16-byte boundary of the padded row,
    //      `off` = tile column of strip x = 0 (4 .. 19).  The clears below run while the copy is in flight. ---------------
    const int pcol0 = x0 + kEdge;                                 // padded column of strip x = 0
    const int tcol0 = (pcol0 - 4) & ~15;
    const int off = pcol0 - tcol0;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (uint32_t)(kTilePitch * (hcell + 6)));
        tma_load_3d(tile, &maps.m[l], &bar, tcol0, y0 - 3 + kEdge, ws.slot0 + b);
    }
    {
        uint4* z = reinterpret_cast<uint4*>(sc);
        for (int i = tid; i < (L.sc_bytes >> 4); i += kFastThreads) z[i] = make_uint4(0, 0, 0, 0);
        uint4* zm = reinterpret_cast<uint4*>(mask);             // mask_words is a multiple of 4 and the array is 16-byte aligned
        for (int i = tid; i < (L.mask_words >> 2); i += kFastThreads) zm[i] = make_uint4(0, 0, 0, 0);
        const unsigned wrecip = (unsigned)g.wcell_recip;
        for (int x = tid; x < gw; x += kFastThreads) {
            xinfo[x] = (uint8_t)(((unsigned)x * wrecip) >> 16);    // cell index of the strip column = x / wcell (exact: x < 256, wcell >= 30)
        }
    }
    mbar_wait(&bar, 0);
    __syncthreads();                                              // clears complete before any warp scores

    const int ini_th = plan->ini_th, lo_th = min(plan->ini_th, plan->min_th);

    // ---- phase 1: compass pretest, 8 pixels per thread (four u16x2 pairs) -> this warp's work list ------------------
    int nwork = 0;
    {
        const unsigned k2 = (unsigned)(0x8000 - 257 - lo_th) * 0x00010001u;
        // aligned octets (8 tile columns = one 64-bit word pair) covering the strip; the warp's rows (py = wid, wid + 8, ...)
        // are walked as ONE stream of octets, 32 per iteration, so that only the last iteration has idle lanes;
        // t / noct by multiplication (t < 512, noct <= 32: exact with a 16-bit reciprocal)
        const int o0 = off >> 3, noct = ((off + gw - 1) >> 3) - o0 + 1;
        const int nrows_w = (ch - wid + kFastWarps - 1) / kFastWarps;
        const int ntask = nrows_w * noct;
        const unsigned recip = 65536u / (unsigned)noct + 1u;
        for (int tb = 0; tb < ntask; tb += 32) {
            const int t = tb + lane;
            const int rl = (int)(((unsigned)t * recip) >> 16);
            const int xq = t - rl * noct;
            const int py = wid + kFastWarps * rl;
            const int xs = 8 * (o0 + xq) - off;                  // strip x of the octet's first pixel (-7 .. gw-1)
            // pass bits: pixel 2i -> bit 2i, pixel 2i+1 -> bit 16 + 2i
            unsigned q = 0;
            if (t < ntask) {
                const uint32_t* rc = tile32 + (py + 3) * kTileW + 2 * (o0 + xq);       // centre-row word pair of the octet
                const uint2 cw = *reinterpret_cast<const uint2*>(rc);
                const unsigned cl = rc[-1], cr = rc[2];
                const uint2 up = *reinterpret_cast<const uint2*>(rc - 3 * kTileW);
                const uint2 dn = *reinterpret_cast<const uint2*>(rc + 3 * kTileW);
                const unsigned lf0 = __funnelshift_r(cl, cw.x, 8), lf1 = __funnelshift_r(cw.x, cw.y, 8);      // x-3 ..
                const unsigned rt0 = __funnelshift_r(cw.x, cw.y, 24), rt1 = __funnelshift_r(cw.y, cr, 24);    // x+3 ..
                const unsigned p0 = pretest_pair(__byte_perm(cw.x, 0, 0x4140), __byte_perm(dn.x, 0, 0x4140), __byte_perm(up.x, 0, 0x4140),
                                                 __byte_perm(rt0, 0, 0x4140), __byte_perm(lf0, 0, 0x4140), k2);
                const unsigned p1 = pretest_pair(__byte_perm(cw.x, 0, 0x4342), __byte_perm(dn.x, 0, 0x4342), __byte_perm(up.x, 0, 0x4342),
                                                 __byte_perm(rt0, 0, 0x4342), __byte_perm(lf0, 0, 0x4342), k2);
                const unsigned p2 = pretest_pair(__byte_perm(cw.y, 0, 0x4140), __byte_perm(dn.y, 0, 0x4140), __byte_perm(up.y, 0, 0x4140),
                                                 __byte_perm(rt1, 0, 0x4140), __byte_perm(lf1, 0, 0x4140), k2);
                const unsigned p3 = pretest_pair(__byte_perm(cw.y, 0, 0x4342), __byte_perm(dn.y, 0, 0x4342), __byte_perm(up.y, 0, 0x4342),
                                                 __byte_perm(rt1, 0, 0x4342), __byte_perm(lf1, 0, 0x4342), k2);
                q = (p0 >> 15) | (p1 >> 13) | (p2 >> 11) | (p3 >> 9);
                if (xs < 0 || xs + 7 >= gw) {                    // first / last octet: drop the pixels outside the strip
                    // pixels k in [max(0, -xs), min(8, gw - xs)) stay; their pass bits sit at k (even k) and 15 + k (odd k)
                    const unsigned m8 = (0xFFu >> (8 - min(8, gw - xs))) & (0xFFu << max(0, -xs));
                    q &= (m8 & 0x55u) | ((m8 & 0xAAu) << 15);
                }
            }
            // append the passing pixels of all lanes: one popcount per lane, warp scan by shuffles, predicated stores
            // (the order of the list is irrelevant: scores go to the score tile, survivors to the row masks)
            const int c = __popc(q);
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up_ = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += up_;
            }
            uint16_t* wp = work + nwork + inc - c;
            nwork += __shfl_sync(0xffffffffu, inc, 31);
            const unsigned ent = (unsigned)((py << 8) + xs);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (q & (1u << ((k & 1) * 16 + (k >> 1) * 2))) *wp++ = (uint16_t)(ent + k);
        }
    }
    __syncwarp();

    // ---- phase 2: exact score of the listed pixels (each warp scores its own list) ---------------------------------
    const uint8_t* t0 = tile + 3 * kTilePitch + off;              // strip pixel (0,0)
    const unsigned lt = (1u << lane) - 1u;
    int ncorner = 0;                                              // corners (m > lo_th) are compacted IN PLACE to work[0 .. ncorner)
    for (int base = 0; base < nwork; base += 32) {
        const int i = base + lane;
        const int e = work[i < nwork ? i : base];
        const uint8_t* c = t0 + e;                                // e = py << 8 | px and the tile pitch is 256
        const unsigned cv = c[0];
        // v = {I - c + 256 (low half), c - I + 256 (high half)}: both halves in [1, 511], no carry between them
        const unsigned K = (256u - cv) + ((cv + 256u) << 16);
        unsigned v[16];
#define FBE_RING(k, dy, dx) v[k] = (unsigned)c[(dy) * kTilePitch + (dx)] * 0xFFFF0001u + K
        FBE_RING(0, 3, 0);   FBE_RING(1, 3, 1);    FBE_RING(2, 2, 2);    FBE_RING(3, 1, 3);
        FBE_RING(4, 0, 3);   FBE_RING(5, -1, 3);   FBE_RING(6, -2, 2);   FBE_RING(7, -3, 1);
        FBE_RING(8, -3, 0);  FBE_RING(9, -3, -1);  FBE_RING(10, -2, -2); FBE_RING(11, -1, -3);
        FBE_RING(12, 0, -3); FBE_RING(13, 1, -3);  FBE_RING(14, 2, -2);  FBE_RING(15, 3, -1);
#undef FBE_RING
        unsigned m3[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(v[k], v[(k + 1) & 15], v[(k + 2) & 15]);
        unsigned M = 0u;                                          // halves are >= 1
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            const unsigned a = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
            const unsigned bq = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
            M = __vimax3_s16x2(M, a, bq);
        }
        const int m = max((int)(M & 0xFFFFu), (int)(M >> 16)) - 256;
        const bool ok = i < nwork && m > lo_th;
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        __syncwarp();                                             // every lane has read its entry before any slot is overwritten
        if (ok) {
            sc[e + xinfo[e & 255] + (kScorePitch + 1)] = (uint8_t)(m - 1);
            work[ncorner + __popc(bal & lt)] = (uint16_t)e;
        }
        ncorner += __popc(bal);
    }
    __syncthreads();                                              // neighbours' scores come from other warps

    // ---- phase 3: strict 8-neighbour NMS inside the cell -> row masks -----------------------------------------------
    uint32_t* mask_ini = mask + ncell * ch * 2;
    __syncwarp();
    for (int i = lane; i < ncorner; i += 32) {
        const int e = work[i];
        const int px = e & 255, py = e >> 8;
        const int cj = xinfo[px];
        const uint8_t* q = sc + e + cj + (kScorePitch + 1);
        const unsigned s = q[0];
        unsigned nb = __vimax3_u32(q[-kScorePitch - 1], q[-kScorePitch], q[-kScorePitch + 1]);
        nb = __vimax3_u32(nb, q[-1], q[1]);
        nb = __vimax3_u32(nb, q[kScorePitch - 1], q[kScorePitch]);
        nb = max(nb, (unsigned)q[kScorePitch + 1]);
        if (s > nb) {
            const int xin = px - cj * wcell;
            const int w = (cj * ch + py) * 2 + (xin >> 5);
            const unsigned bit = 1u << (xin & 31);
            atomicOr(&mask[w], bit);
            if ((int)s >= ini_th) atomicOr(&mask_ini[w], bit);
        }
    }
    __syncthreads();

    // ---- phase 4: ordered emission, one warp per cell ---------------------------------------------------------------
    for (int cj = wid; cj < ncell; cj += kFastWarps) {
        const uint32_t* ml = mask + cj * ch * 2;
        const uint32_t* mi = mask_ini + cj * ch * 2;
        const int r0 = lane, r1 = lane + 32;
        unsigned long long a0 = 0, a1 = 0, b0 = 0, b1 = 0;
        if (r0 < ch) { a0 = ml[r0 * 2] | ((unsigned long long)ml[r0 * 2 + 1] << 32); b0 = mi[r0 * 2] | ((unsigned long long)mi[r0 * 2 + 1] << 32); }
        if (r1 < ch) { a1 = ml[r1 * 2] | ((unsigned long long)ml[r1 * 2 + 1] << 32); b1 = mi[r1 * 2] | ((unsigned long long)mi[r1 * 2 + 1] << 32); }
        if (__any_sync(0xffffffffu, (b0 | b1) != 0ull)) { a0 = b0; a1 = b1; }
        const int c0 = __popcll(a0), c1 = __popcll(a1);
        const int s01 = warp_incl_scan(c0 | (c1 << 16), lane);          // both row halves in one scan (a cell holds < 2^16 survivors)
        const int t01 = __shfl_sync(0xffffffffu, s01, 31);
        const int s0 = s01 & 0xFFFF, s1 = s01 >> 16, tot0 = t01 & 0xFFFF, tot1 = t01 >> 16;
        int o0 = s0 - c0, o1 = tot0 + s1 - c1;
        uint32_t* slots = ws.slots + (size_t)b * plan->slots_total + g.slot_base + (size_t)(ci * ncols + cj0 + cj) * g.cell_cap;
        const int cx = cj * wcell;
        const uint8_t* scc = sc + cx + cj + 1;                   // score column of the cell's x = 0
        while (a0) {
            const int x = __ffsll((long long)a0) - 1;
            a0 &= a0 - 1;
            slots[o0++] = pack_key(x0 + cx + x, y0 + r0, scc[(r0 + 1) * kScorePitch + x]);
        }
        while (a1) {
            const int x = __ffsll((long long)a1) - 1;
            a1 &= a1 - 1;
            slots[o1++] = pack_key(x0 + cx + x, y0 + r1, scc[(r1 + 1) * kScorePitch + x]);
        }
        if (lane == 0) count_out[cj] = tot0 + tot1;
    }
}

int launch_fast_cells(const Plan& hp, const Plan* dp, const Workspace& ws, const TmaMaps& maps, int nimg, cudaStream_t st) {
    size_t smem = 0;
    for (int l = 0; l < hp.nlevels; ++l) {
        const LevelGeom& g = hp.lv[l];
        if (g.gcells > kFastMaxCells) { set_error("FAST strip with more than 7 cells (cells are at least 30 px wide)"); return FBE_E_UNSUPPORTED; }
        smem = std::max(smem, (size_t)fast_layout(g.gcells * g.wcell, g.hcell, g.hcell, g.gcells).total);
    }
    if (smem > 200 * 1024) { set_error("FAST strip too large for shared memory"); return FBE_E_UNSUPPORTED; }
    if (smem > 48 * 1024) FBE_CUDA(cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(hp.ngroups_total, nimg);
    FBE_CUDA(launch_dep(k_fast_cells, grid, dim3(kFastThreads), smem, st, dp, ws, maps));
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
