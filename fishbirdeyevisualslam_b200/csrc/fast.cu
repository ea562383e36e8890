// Per-cell FAST-9/16 detection with threshold fallback (ComputeKeyPointsOctTree cell loop,
// src/ORBextractor.cc:765-829) as ONE kernel: a CTA owns one ~30x30 FAST cell of one level of one image.
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
// Output per cell: survivors in (y, x) order as packed keys in the cell's private slot range + a count.  The octree
// kernel concatenates cells in row-major order, which reproduces vToDistributeKeys order.
#include "fbe_internal.cuh"

namespace fbe {

constexpr int kFastThreads = 256;

// m for one pixel given its 16 ring differences d[k] = I_k - c (k clockwise).
__device__ __forceinline__ int fast_m_from_ring(const int (&d)[16]) {
    // sliding-window minimum (bright) / maximum (dark) of length 9 over the circular array, by doubling
    int lo1[16], hi1[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { lo1[k] = min(d[k], d[(k + 1) & 15]); hi1[k] = max(d[k], d[(k + 1) & 15]); }
    int lo2[16], hi2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { lo2[k] = min(lo1[k], lo1[(k + 2) & 15]); hi2[k] = max(hi1[k], hi1[(k + 2) & 15]); }
    int lo4[16], hi4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { lo4[k] = min(lo2[k], lo2[(k + 4) & 15]); hi4[k] = max(hi2[k], hi2[(k + 4) & 15]); }
    // max over arcs of (-hi9) == -(min over arcs of hi9).  NB: the direct form max(best, max(lo9, -hi9)) is
    // MISCOMPILED by nvcc 12.9 for sm_100a (the negation is dropped when the chain is fused into 3-input VIMNMX;
    // reproduced in isolation on a B200, see DESIGN.md "toolchain notes"), so the negation is hoisted out of the chain.
    int maxlo = -256, minhi = 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        maxlo = max(maxlo, min(lo4[k], d[(k + 8) & 15]));     // lo4[k] covers k..k+7
        minhi = min(minhi, max(hi4[k], d[(k + 8) & 15]));
    }
    return max(0, max(maxlo, -minhi));
}

__global__ void __launch_bounds__(kFastThreads) k_fast_cells(const Plan* __restrict__ plan, Workspace ws) {
    extern __shared__ uint8_t smem[];
    __shared__ int s_warp[kFastThreads / 32];
    __shared__ int s_ini_count;
    __shared__ int s_running;

    const int b = blockIdx.y;
    int cell = blockIdx.x;
    // locate the level of this cell
    int l = 0;
    const int nl = plan->nlevels;
    while (l + 1 < nl && cell >= plan->lv[l + 1].cell_base) ++l;
    const LevelGeom g = plan->lv[l];
    cell -= g.cell_base;
    const int ci = cell / g.ncols, cj = cell - ci * g.ncols;

    // detection region of the cell in level coordinates
    const int x0 = kEdge + cj * g.wcell, y0 = kEdge + ci * g.hcell;
    const int x1 = min(x0 + g.wcell, g.w - kEdge), y1 = min(y0 + g.hcell, g.h - kEdge);
    const int cw = x1 - x0, ch = y1 - y0;
    int* count_out = ws.cell_count + (size_t)b * plan->ncells_total + g.cell_base + cell;
    if (cw <= 0 || ch <= 0) {
        if (threadIdx.x == 0) *count_out = 0;
        return;
    }
    const int tw = cw + 6, th_ = ch + 6;            // staged tile incl. 3-px ring halo
    const int tpitch = (tw + 3) & ~3;
    const int spitch = cw + 2;                      // score tile with a 1-px zero margin
    uint8_t* tile = smem;
    uint8_t* sc = smem + ((tpitch * th_ + 15) & ~15);
    uint8_t* sv = sc + (((cw + 2) * (ch + 2) + 15) & ~15);   // survivor scores (0 = not a survivor)

    const uint8_t* img = ws.pyr + (size_t)b * plan->pyr_bytes + g.img_off;
    // level (x,y) lives at padded (x+19, y+19)
    const uint8_t* src = img + (size_t)(y0 - 3 + kEdge) * g.pitch + (x0 - 3 + kEdge);
    for (int i = threadIdx.x; i < tw * th_; i += kFastThreads) {
        int ty = i / tw, tx = i - ty * tw;
        tile[ty * tpitch + tx] = src[(size_t)ty * g.pitch + tx];
    }
    for (int i = threadIdx.x; i < (cw + 2) * (ch + 2); i += kFastThreads) sc[i] = 0;
    if (threadIdx.x == 0) { s_ini_count = 0; s_running = 0; }
    __syncthreads();

    const int lo_th = min(plan->ini_th, plan->min_th);
    const int npix = cw * ch;
    for (int p = threadIdx.x; p < npix; p += kFastThreads) {
        const int py = p / cw, px = p - py * cw;
        const uint8_t* c = tile + (py + 3) * tpitch + (px + 3);
        const int cv = c[0];
        // exact necessary condition: every 9-arc contains k or k+8 for each k
        const int d0 = c[3 * tpitch] - cv, d8 = c[-3 * tpitch] - cv;
        const int d4 = c[3] - cv, d12 = c[-3] - cv;
        bool br = (d0 > lo_th || d8 > lo_th) && (d4 > lo_th || d12 > lo_th);
        bool dk = (d0 < -lo_th || d8 < -lo_th) && (d4 < -lo_th || d12 < -lo_th);
        int score = 0;
        if (br || dk) {
            int d[16];
            d[0] = d0; d[4] = d4; d[8] = d8; d[12] = d12;
            d[1] = c[3 * tpitch + 1] - cv;  d[2] = c[2 * tpitch + 2] - cv;  d[3] = c[tpitch + 3] - cv;
            d[5] = c[-tpitch + 3] - cv;     d[6] = c[-2 * tpitch + 2] - cv; d[7] = c[-3 * tpitch + 1] - cv;
            d[9] = c[-3 * tpitch - 1] - cv; d[10] = c[-2 * tpitch - 2] - cv; d[11] = c[-tpitch - 3] - cv;
            d[13] = c[tpitch - 3] - cv;     d[14] = c[2 * tpitch - 2] - cv; d[15] = c[3 * tpitch - 1] - cv;
            const int m = fast_m_from_ring(d);
            if (m > lo_th) score = m - 1;
        }
        sc[(py + 1) * spitch + (px + 1)] = (uint8_t)score;
    }
    __syncthreads();

    // strict 8-neighbour NMS inside the cell
    int my_ini = 0;
    for (int p = threadIdx.x; p < npix; p += kFastThreads) {
        const int py = p / cw, px = p - py * cw;
        const uint8_t* q = sc + (py + 1) * spitch + (px + 1);
        const int s = q[0];
        int keep = 0;
        if (s > 0) {
            int nb = max(max(max(q[-spitch - 1], q[-spitch]), max(q[-spitch + 1], q[-1])),
                         max(max(q[1], q[spitch - 1]), max(q[spitch], q[spitch + 1])));
            keep = s > nb;
        }
        sv[p] = keep ? (uint8_t)s : 0;
        my_ini += (keep && s >= plan->ini_th);
    }
    my_ini = __reduce_add_sync(0xffffffffu, my_ini);
    if ((threadIdx.x & 31) == 0 && my_ini) atomicAdd(&s_ini_count, my_ini);
    __syncthreads();
    const int emit_th = s_ini_count > 0 ? plan->ini_th : plan->min_th;

    // ordered (y,x) compaction into the cell's slots
    uint32_t* slots = ws.slots + (size_t)b * plan->slots_total + g.slot_base + (size_t)cell * g.cell_cap;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < npix; base += kFastThreads) {
        const int p = base + threadIdx.x;
        int s = 0;
        if (p < npix) s = sv[p];
        const bool flag = s >= emit_th && s > 0;
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kFastThreads / 32; ++w) {
            int c = s_warp[w];
            if (w < wid) before += c;
            total += c;
        }
        const int run = s_running;
        if (flag) {
            const int py = p / cw, px = p - py * cw;
            slots[run + before + __popc(bal & ((1u << lane) - 1))] = pack_key(x0 + px, y0 + py, s);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_running = run + total;
        // s_running is re-read only after the next __syncthreads in the following iteration
    }
    __syncthreads();
    if (threadIdx.x == 0) *count_out = s_running;
}

int launch_fast_cells(const Plan& hp, const Plan* dp, const Workspace& ws, int nimg, cudaStream_t st) {
    const int tw = hp.max_cell_w + 6, th = hp.max_cell_h + 6;
    const int tpitch = (tw + 3) & ~3;
    size_t smem = ((tpitch * th + 15) & ~15) + (((hp.max_cell_w + 2) * (hp.max_cell_h + 2) + 15) & ~15) +
                  ((hp.max_cell_w * hp.max_cell_h + 15) & ~15);
    if (smem > 200 * 1024) { set_error("FAST cell too large for shared memory"); return FBE_E_UNSUPPORTED; }
    if (smem > 48 * 1024) FBE_CUDA(cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(hp.ncells_total, nimg);
    k_fast_cells<<<grid, kFastThreads, smem, st>>>(dp, ws);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
