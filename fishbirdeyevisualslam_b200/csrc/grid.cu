// Frame::AssignFeaturesToGrid / PosInGrid / PosInGridBirdview (src/Frame.cc:381-411, 548-570) as a CSR build.
// cell(ix,iy) = (round((x-minX)*invW), round((y-minY)*invH)) -- `round`, not floor (quirk Q1) -- keypoints whose cell
// falls outside the grid are dropped.  mGrid[ix][iy] keeps keypoint indices in ascending order; the CSR uses the same
// [ix*grows + iy] ordering so that "ix outer, iy inner, in-cell order" traversal (GetFeaturesInArea) is a linear walk.
// One CTA per frame: shared-memory histogram, block scan, atomic placement, then each cell's few items are sorted so the
// result is the stable (index-ordered) bucket the reference builds with push_back.
#include "fbe_internal.cuh"

namespace fbe {

constexpr int kGridThreads = 1024;

__global__ void __launch_bounds__(kGridThreads) k_grid_build(const fbe_keypoint* __restrict__ kps, const int* __restrict__ n_arr,
                                                             int stride, float min_x, float min_y, float inv_w, float inv_h,
                                                             int gcols, int grows, int* __restrict__ cell_of,
                                                             int* __restrict__ start_out, int* __restrict__ items_out) {
    extern __shared__ int s_cnt[];          // gcells + 1
    __shared__ int s_warp[kGridThreads / 32];
    __shared__ int s_carry;
    pdl_launch_dependents();
    pdl_wait();
    const int f = blockIdx.x;
    const int n = n_arr[f];
    const int gcells = gcols * grows;
    const fbe_keypoint* kp = kps + (size_t)f * stride;
    int* cof = cell_of + (size_t)f * stride;
    int* start = start_out + (size_t)f * (gcells + 1);
    int* items = items_out + (size_t)f * stride;

    for (int i = threadIdx.x; i <= gcells; i += kGridThreads) s_cnt[i] = 0;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kGridThreads) {
        const float px = roundf(__fmul_rn(__fsub_rn(kp[i].x, min_x), inv_w));
        const float py = roundf(__fmul_rn(__fsub_rn(kp[i].y, min_y), inv_h));
        int cell = -1;
        if (px >= 0.f && px < (float)gcols && py >= 0.f && py < (float)grows) cell = (int)px * grows + (int)py;
        cof[i] = cell;
        if (cell >= 0) atomicAdd(&s_cnt[cell], 1);
    }
    __syncthreads();
    // exclusive scan of s_cnt[0..gcells) in place -> cell starts; s_cnt[gcells] = total
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < gcells; base += kGridThreads) {
        const int i = base + threadIdx.x;
        const int v = i < gcells ? s_cnt[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kGridThreads / 32; ++w) {
            int c = s_warp[w];
            if (w < wid) before += c;
            tot += c;
        }
        const int carry = s_carry;
        if (i < gcells) { s_cnt[i] = carry + before + inc - v; start[i] = carry + before + inc - v; }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[gcells] = s_carry;
    __syncthreads();
    // placement (order inside a cell is fixed up below); s_cnt now acts as the per-cell cursor
    for (int i = threadIdx.x; i < n; i += kGridThreads) {
        const int cell = cof[i];
        if (cell >= 0) items[atomicAdd(&s_cnt[cell], 1)] = i;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < gcells; c += kGridThreads) {
        const int s = start[c], e = s_cnt[c];     // cursor ended at the cell's end
        for (int i = s + 1; i < e; ++i) {
            const int v = items[i];
            int j = i - 1;
            while (j >= s && items[j] > v) { items[j + 1] = items[j]; --j; }
            items[j + 1] = v;
        }
    }
}

int launch_grid_build(const fbe_keypoint* d_kps, const int* d_n, int n_stride, int nframes, float min_x, float min_y,
                      float inv_w, float inv_h, int gcols, int grows, int* d_cell_of, int* d_start, int* d_items,
                      cudaStream_t st) {
    const size_t smem = (size_t)(gcols * grows + 1) * sizeof(int);
    if (smem > 48 * 1024) { set_error("grid too large"); return FBE_E_UNSUPPORTED; }
    FBE_CUDA(launch_dep(k_grid_build, dim3(nframes), dim3(kGridThreads), smem, st, d_kps, d_n, n_stride, min_x, min_y, inv_w, inv_h, gcols, grows,
                        d_cell_of, d_start, d_items));
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int launch_grid(const Plan& hp, const Plan* dp, const Workspace& ws, int nimg, cudaStream_t st) {
    (void)dp;
    if (hp.grid_cols <= 0) return FBE_OK;
    return launch_grid_build(ws.out_kps_un, ws.out_n, hp.kp_cap_total, nimg, hp.grid_min_x, hp.grid_min_y, hp.grid_inv_w,
                             hp.grid_inv_h, hp.grid_cols, hp.grid_rows, ws.out_cell, ws.grid_start, ws.grid_items, st);
}

}  // namespace fbe
