// The bird-view feature path the reference actually ships (src/Frame.cc:336-355):
//     cv::Ptr<cv::ORB> extractorBird = cv::ORB::create(2000);
//     extractorBird->detect(mBirdviewImg, preKeysBird, mBirdviewMask);      -> fbe_bird_orb_detect
//     GuidenceKeyBirdPts(preKeysBird); cv::cornerSubPix(...)                 -> bird_refine.cu
//     extractorBird->compute(mBirdviewImg, mvKeysBird, mDescriptorsBird);   -> fbe_bird_orb_compute
// and the whole block as one device-resident call (fbe_bird_features).  cv::ORB is OpenCV code (modules/features2d/src/orb.cpp)
// with its own pyramid, scoring and ordering; everything here follows the oracle restatement oracle/cvorb_oracle.cpp, which is
// pinned to cv2 4.13.0 byte for byte (keypoint order included).  Kernels, all batched over frames (blockIdx.y or .z = frame):
//   k_borb_level0 / k_borb_resize   image + mask pyramid: level l is resized from level l-1 with INTER_LINEAR_EXACT (8.8
//                                   fixed-point taps, (.. + 32768) >> 16), each level stored with its 32-px REFLECT_101 frame;
//                                   the mask is binarised (cv2 4.13: any non-zero value keeps) and re-thresholded per level.
//   k_borb_fast                     whole-level FAST-9/16 (threshold 20) + 3x3 NMS + mask + 31-px border filter, one CTA per band
//                                   of 8 rows: packed s16x2 DPX scores in shared memory, survivors emitted in raster order.
//   k_borb_gather                   bands -> per-level candidate lists (raster order = cv::FAST's output order).
//   k_borb_retain                   KeyPointsFilter::retainBest: std::nth_element + std::partition REPLAYED in parallel --
//                                   libstdc++'s introselect (median of three to the front, unguarded Hoare partition, insertion
//                                   sort of the last <= 3) leaves the survivors in a specific order that cv::ORB's output order
//                                   (and so every downstream index of the reference) inherits.  A Hoare partition pass is
//                                   order-equivalent to: swap the k-th element from the left that stops the left scan with the
//                                   k-th from the right that stops the right scan, for all k while left < right -- two scans
//                                   and one scatter per pass instead of a pointer chase.  When introselect runs out of its depth
//                                   budget (adversarial inputs only) libstdc++ finishes with __heap_select + iter_swap; that exit
//                                   is replayed as well, serially (tests/test_gpu_bird_orb.py drives it with McIlroy's adversary).
//   k_borb_harris                   HarrisResponses (7x7 block of Sobel-like sums, fp32 formula without contraction).
//   k_borb_finish                   second retainBest result -> ICAngles (integer moments, fastAtan2) -> keypoint records.
//   k_borb_prefilter                compute(): runByImageBorder on the rounded positions + regrouping by octave when unsorted.
//   k_borb_blur                     the Gaussian cv::ORB ends up with on a pyramid level: NOT the bit-exact 8-bit path (the level
//                                   is a submatrix) but sepFilter2D with the float kernel, FMA-contracted like OpenCV's AVX2
//                                   build: row taps left to right, column taps centre outwards, cvRound.
//   k_borb_describe                 steered BRIEF: one warp per keypoint, lane = descriptor byte.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>
#include "fbe_internal.cuh"
#include "orb_device.cuh"

namespace fbe {

constexpr int kBorbBorder = 32, kBorbEdge = 31, kBorbLevels = 8, kBorbFastTh = 20, kBorbBand = 8, kBorbThreads = 256;

__constant__ int8_t c_borb_pattern[256 * 4] = {
#include "orb_pattern.inc"
};

// row half-widths of the radius-15 disc (computeKeyPoints' umax: the same table as ORBextractor's, orb_device.cuh)
__constant__ int c_borb_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

struct BorbTab { short o0, o1, c0, c1; };          // exact-resize taps of one destination coordinate: c0 * S[o0] + c1 * S[o1], c0 + c1 = 256

struct BorbLevel {
    int w, h, pitch, ph;        // ROI size, padded row pitch (multiple of 16) and padded rows (h + 64)
    int img_off;                // byte offset of the padded level in the per-frame pyramid slab
    int mask_off;               // byte offset of the (unpadded, pitch w) mask level in the per-frame mask slab
    float scale;                // layerScale = (float)pow(1.2, level)
    int nfeat;                  // nfeaturesPerLevel
    int kw, kh;                 // keypoint region (w - 62, h - 62), <= 0: no keypoints on this level
    int nbands, band_off;       // FAST bands of 8 rows; first band in the per-frame band numbering
    int band_cap;               // survivors per band
    int cand_cap, cand_off;     // candidates of the level
    int tabx_off, taby_off;     // resize taps (levels >= 1)
    int blur_tiles;             // 64 x 32 tiles of the blur kernel
};

struct BorbPlan {
    int rows, cols, nfeatures;
    int pyr_bytes, mask_bytes, bands_total, cand_total, band_slots_total;
    BorbLevel lv[kBorbLevels];
};

__host__ __device__ inline int borb_reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// ---- pyramid -------------------------------------------------------------------------------------------------------------------
__global__ void k_borb_level0(const BorbPlan* __restrict__ plan, const uint8_t* __restrict__ imgs, size_t step, size_t stride,
                              const uint8_t* __restrict__ masks, size_t mstep, size_t mstride, uint8_t* __restrict__ pyr,
                              uint8_t* __restrict__ mpyr) {
    const BorbLevel& g = plan->lv[0];
    const int b = blockIdx.z, px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y;
    if (px >= g.w + 2 * kBorbBorder) return;
    const int x = borb_reflect101(px - kBorbBorder, g.w), y = borb_reflect101(py - kBorbBorder, g.h);
    pyr[(size_t)b * plan->pyr_bytes + g.img_off + (size_t)py * g.pitch + px] = imgs[(size_t)b * stride + (size_t)y * step + x];
    if (masks && px >= kBorbBorder && px < g.w + kBorbBorder && py >= kBorbBorder && py < g.h + kBorbBorder)
        mpyr[(size_t)b * plan->mask_bytes + g.mask_off + (size_t)(py - kBorbBorder) * g.w + (px - kBorbBorder)] =
            masks[(size_t)b * mstride + (size_t)(py - kBorbBorder) * mstep + (px - kBorbBorder)] ? 255 : 0;
}

// level l from the ROI of level l-1: the image padded (frame = reflection of the level itself) and, when there is a mask, the mask
// level unpadded and re-thresholded (> 254 kept), in the same launch (blockIdx.y < ph: image rows, above: mask rows)
__device__ __forceinline__ unsigned borb_exact_tap(const uint8_t* S, int sp, const BorbTab tx, const BorbTab ty) {
    const uint8_t* r0 = S + (size_t)ty.o0 * sp;
    const uint8_t* r1 = S + (size_t)ty.o1 * sp;
    const unsigned h0 = (unsigned)tx.c0 * r0[tx.o0] + (unsigned)tx.c1 * r0[tx.o1];
    const unsigned h1 = (unsigned)tx.c0 * r1[tx.o0] + (unsigned)tx.c1 * r1[tx.o1];
    return ((unsigned)ty.c0 * h0 + (unsigned)ty.c1 * h1 + 32768u) >> 16;
}

__global__ void k_borb_resize(const BorbPlan* __restrict__ plan, const BorbTab* __restrict__ tabs, int l, uint8_t* __restrict__ pyr,
                              uint8_t* __restrict__ mpyr) {
    const BorbLevel& g = plan->lv[l];
    const BorbLevel& s = plan->lv[l - 1];
    const int b = blockIdx.z, px = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)blockIdx.y < g.ph) {
        const int py = blockIdx.y;
        if (px >= g.w + 2 * kBorbBorder) return;
        const int x = borb_reflect101(px - kBorbBorder, g.w), y = borb_reflect101(py - kBorbBorder, g.h);
        const uint8_t* S = pyr + (size_t)b * plan->pyr_bytes + s.img_off + (size_t)kBorbBorder * s.pitch + kBorbBorder;
        pyr[(size_t)b * plan->pyr_bytes + g.img_off + (size_t)py * g.pitch + px] =
            (uint8_t)borb_exact_tap(S, s.pitch, tabs[g.tabx_off + x], tabs[g.taby_off + y]);
    } else {
        const int py = blockIdx.y - g.ph;
        if (px >= g.w) return;
        const uint8_t* S = mpyr + (size_t)b * plan->mask_bytes + s.mask_off;
        const unsigned v = borb_exact_tap(S, s.w, tabs[g.tabx_off + px], tabs[g.taby_off + py]);
        mpyr[(size_t)b * plan->mask_bytes + g.mask_off + (size_t)py * g.w + px] = (uint8_t)(v > 254u ? v : 0u);
    }
}

// ---- FAST ----------------------------------------------------------------------------------------------------------------------
// m = max over the 16 arcs of 9 contiguous ring pixels of min |I - c| with one sign (corner at threshold t <=> m > t, score m - 1);
// the compass pretest (two adjacent compass points must agree) rejects most pixels before the packed s16x2 DPX network
__device__ __forceinline__ int borb_fast_m(const uint8_t* c, int p, int th) {
    const int cv = c[0];
    const int i0 = c[3 * p], i8 = c[-3 * p], i4 = c[3], i12 = c[-3];
    const int e = min(max(i0, i8), max(i4, i12)), f = max(min(i0, i8), min(i4, i12));
    if (e - cv <= th && cv - f <= th) return 0;
    const unsigned K = (256u - (unsigned)cv) + (((unsigned)cv + 256u) << 16);
    unsigned v[16];
#define FBE_RING(k, dy, dx) v[k] = (unsigned)c[(dy) * p + (dx)] * 0xFFFF0001u + K
    FBE_RING(0, 3, 0);   FBE_RING(1, 3, 1);    FBE_RING(2, 2, 2);    FBE_RING(3, 1, 3);
    FBE_RING(4, 0, 3);   FBE_RING(5, -1, 3);   FBE_RING(6, -2, 2);   FBE_RING(7, -3, 1);
    FBE_RING(8, -3, 0);  FBE_RING(9, -3, -1);  FBE_RING(10, -2, -2); FBE_RING(11, -1, -3);
    FBE_RING(12, 0, -3); FBE_RING(13, 1, -3);  FBE_RING(14, 2, -2);  FBE_RING(15, 3, -1);
#undef FBE_RING
    unsigned m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(v[k], v[(k + 1) & 15], v[(k + 2) & 15]);
    unsigned M = 0u;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned a = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
        const unsigned bq = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
        M = __vimax3_s16x2(M, a, bq);
    }
    return max((int)(M & 0xFFFFu), (int)(M >> 16)) - 256;
}

// one CTA = one band of 8 keypoint rows of one level of one frame.  Keypoint region K = [31, w-31) x [31, h-31) (the detector's
// runByImageBorder); scores are needed on K plus a 1-px ring for the NMS.
__global__ void __launch_bounds__(kBorbThreads) k_borb_fast(const BorbPlan* __restrict__ plan, const short* __restrict__ band_level,
                                                            const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ mpyr, bool has_mask,
                                                            uint2* __restrict__ band_slots, int* __restrict__ band_count) {
    extern __shared__ uint8_t s_borb[];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int l = band_level[blockIdx.x];
    const BorbLevel& g = plan->lv[l];
    const int band = blockIdx.x - g.band_off;
    const int sw = g.kw + 2;                                     // score columns: x = 30 .. w - 31
    uint8_t* sc = s_borb;                                        // [kBorbBand + 2][sw]
    uint32_t* rowmask = reinterpret_cast<uint32_t*>(s_borb + (((kBorbBand + 2) * sw + 15) & ~15));    // [kBorbBand][words]
    const int words = (g.kw + 31) >> 5;
    __shared__ int row_cnt[kBorbBand], row_off[kBorbBand + 1];
    const int y0 = kBorbEdge + band * kBorbBand;                 // first keypoint row of the band
    const int nrows = min(kBorbBand, g.h - kBorbEdge - y0);
    const uint8_t* img = pyr + (size_t)b * plan->pyr_bytes + g.img_off + (size_t)kBorbBorder * g.pitch + kBorbBorder;   // ROI (0, 0)
    // scores of rows y0 - 1 .. y0 + nrows, columns 30 .. w - 31
    for (int i = tid; i < (nrows + 2) * sw; i += kBorbThreads) {
        const int r = i / sw, cx = i - r * sw;
        const int m = borb_fast_m(img + (size_t)(y0 - 1 + r) * g.pitch + (kBorbEdge - 1 + cx), g.pitch, kBorbFastTh);
        sc[i] = (uint8_t)(m > kBorbFastTh ? m - 1 : 0);
    }
    for (int i = tid; i < kBorbBand * words; i += kBorbThreads) rowmask[i] = 0u;
    __syncthreads();
    // strict 3x3 non-maximum suppression + mask, one warp per row, 32 columns per step
    const uint8_t* mk = has_mask ? mpyr + (size_t)b * plan->mask_bytes + g.mask_off : nullptr;
    for (int r = wid; r < nrows; r += kBorbThreads / 32) {
        int cnt = 0;
        for (int w0 = 0; w0 < words; ++w0) {
            const int kx = w0 * 32 + lane;                       // keypoint column index: x = 31 + kx
            bool keep = false;
            if (kx < g.kw) {
                const uint8_t* q = sc + (r + 1) * sw + kx + 1;
                const int s = q[0];
                if (s > 0) {
                    int nb = max(max((int)q[-sw - 1], (int)q[-sw]), (int)q[-sw + 1]);
                    nb = max(nb, max((int)q[-1], (int)q[1]));
                    nb = max(nb, max(max((int)q[sw - 1], (int)q[sw]), (int)q[sw + 1]));
                    keep = s > nb && (!mk || mk[(size_t)(y0 + r) * g.w + (kBorbEdge + kx)] != 0);
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) rowmask[r * words + w0] = bal;
            cnt += __popc(bal);
        }
        if (lane == 0) row_cnt[r] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        for (int r = 0; r < nrows; ++r) { row_off[r] = o; o += row_cnt[r]; }
        row_off[nrows] = o;
        band_count[(size_t)b * plan->bands_total + blockIdx.x] = o;
    }
    __syncthreads();
    uint2* out = band_slots + (size_t)b * plan->band_slots_total + (size_t)blockIdx.x * plan->lv[0].band_cap;   // uniform stride: widest level
    for (int r = wid; r < nrows; r += kBorbThreads / 32) {
        int o = row_off[r];
        for (int w0 = 0; w0 < words; ++w0) {
            const unsigned bal = rowmask[r * words + w0];
            if (bal & (1u << lane)) {
                const int kx = w0 * 32 + lane;
                out[o + __popc(bal & ((1u << lane) - 1u))] = make_uint2((unsigned)(kBorbEdge + kx) | ((unsigned)(y0 + r) << 16),
                                                                         (unsigned)sc[(r + 1) * sw + kx + 1]);
            }
            o += __popc(bal);
        }
    }
}

// bands of a level -> the level's candidate list, in raster order (cv::FAST's output order)
__global__ void __launch_bounds__(kBorbThreads) k_borb_gather(const BorbPlan* __restrict__ plan, const uint2* __restrict__ band_slots,
                                                              const int* __restrict__ band_count, uint32_t* __restrict__ cand_xy,
                                                              float* __restrict__ cand_resp, int* __restrict__ cand_n) {
    const int l = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const BorbLevel& g = plan->lv[l];
    __shared__ int off[512], cnt[512];
    const int* bc = band_count + (size_t)b * plan->bands_total + g.band_off;
    if (tid < 32) {                                              // exclusive prefix of the band counts, 32 bands per step
        int carry = 0;
        for (int c0 = 0; c0 < g.nbands; c0 += 32) {
            const int v = c0 + tid < g.nbands ? bc[c0 + tid] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
            if (c0 + tid < g.nbands) { off[c0 + tid] = carry + inc - v; cnt[c0 + tid] = v; }
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (tid == 0) cand_n[b * kBorbLevels + l] = carry;
    }
    __syncthreads();
    // one warp per band (counts and offsets come from shared memory, so the bands' loads are independent of each other)
    uint32_t* oxy = cand_xy + (size_t)b * plan->cand_total + g.cand_off;
    float* oresp = cand_resp + (size_t)b * plan->cand_total + g.cand_off;
    for (int i = tid >> 5; i < g.nbands; i += kBorbThreads / 32) {
        const uint2* src = band_slots + (size_t)b * plan->band_slots_total + (size_t)(g.band_off + i) * plan->lv[0].band_cap;
        const int n = cnt[i], o = off[i];
        for (int k = tid & 31; k < n; k += 32) {
            const uint2 v = src[k];
            oxy[o + k] = v.x;
            oresp[o + k] = (float)v.y;
        }
    }
}

// ---- KeyPointsFilter::retainBest, replayed ---------------------------------------------------------------------------------------
__device__ __forceinline__ int block_sum(int v, int* s) {        // sum over the CTA, result in every thread
    const int tid = threadIdx.x;
    v = __reduce_add_sync(0xffffffffu, v);
    __syncthreads();
    if ((tid & 31) == 0) s[tid >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int i = 0; i < kBorbThreads / 32; ++i) t += s[i];
    return t;
}

// exclusive prefix (LEFT = true) or exclusive suffix (LEFT = false) of one int per thread; total in every thread
template <bool LEFT>
__device__ __forceinline__ int block_excl(int v, int* s, int& total) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = LEFT ? __shfl_up_sync(0xffffffffu, inc, o) : __shfl_down_sync(0xffffffffu, inc, o);
        if (LEFT ? lane >= o : lane + o < 32) inc += t;
    }
    __syncthreads();
    if (lane == (LEFT ? 31 : 0)) s[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kBorbThreads / 32; ++i) {
        const int x = s[i];
        tot += x;
        if (LEFT ? i < w : i > w) base += x;
    }
    total = tot;
    return base + inc - v;
}

// One pass of the swap-pairing that both libstdc++ partitions reduce to, over [lo, hi): element i stops the left scan when
// stopL(key[i]) and the right scan when stopR(key[i]).  Returns K (number of swaps); lpos / rpos hold the stoppers by rank and
// totalL / totalR their counts.  key / idx are permuted in place.
template <class SL, class SR>
__device__ int swap_pairs(float* key, int* idx, int* lpos, int* rpos, int lo, int hi, SL stopL, SR stopR, int* s, int& totalL, int& totalR) {
    const int tid = threadIdx.x, len = hi - lo;
    const int chunk = (len + kBorbThreads - 1) / kBorbThreads;
    const int i0 = min(lo + tid * chunk, hi), i1 = min(i0 + chunk, hi);
    int cl = 0, cr = 0;
    for (int i = i0; i < i1; ++i) { const float v = key[i]; cl += stopL(v) ? 1 : 0; cr += stopR(v) ? 1 : 0; }
    int bl = block_excl<true>(cl, s, totalL);
    int br = block_excl<false>(cr, s, totalR);
    for (int i = i0; i < i1; ++i) if (stopL(key[i])) lpos[bl++] = i;
    for (int i = i1 - 1; i >= i0; --i) if (stopR(key[i])) rpos[br++] = i;
    __syncthreads();
    const int m = min(totalL, totalR);
    int c = 0;
    for (int k = tid; k < m; k += kBorbThreads) c += lpos[k] < rpos[k] ? 1 : 0;       // monotone in k: true exactly for k < K
    const int K = block_sum(c, s);
    for (int k = tid; k < K; k += kBorbThreads) {
        const int a = lpos[k], b = rpos[k];
        const float ka = key[a]; key[a] = key[b]; key[b] = ka;
        const int ia = idx[a]; idx[a] = idx[b]; idx[b] = ia;
    }
    __syncthreads();
    return K;
}

// libstdc++'s heap algorithms on (key, idx) pairs with comp(a, b) = a.key > b.key, as one thread (bits/stl_heap.h: __adjust_heap with
// its trailing __push_heap, __make_heap, __pop_heap, __heap_select).  Only reached when introselect runs out of its depth budget,
// which takes an adversarial input; the replay is serial like the original.
__device__ void adjust_heap_serial(float* key, int* idx, int hole, int len, float vk, int vi) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (key[child] > key[child - 1]) --child;
        key[hole] = key[child]; idx[hole] = idx[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        key[hole] = key[child - 1]; idx[hole] = idx[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && key[parent] > vk) {
        key[hole] = key[parent]; idx[hole] = idx[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    key[hole] = vk; idx[hole] = vi;
}

// std::__heap_select(first, middle, last, greater) followed by std::iter_swap(first, nth) -- the depth-limit exit of std::__introselect
__device__ void heap_select_serial(float* key, int* idx, int first, int middle, int last, int nth) {
    float* k = key + first;
    int* x = idx + first;
    const int len = middle - first;
    if (len >= 2) {
        for (int parent = (len - 2) / 2;; --parent) {
            adjust_heap_serial(k, x, parent, len, k[parent], x[parent]);
            if (parent == 0) break;
        }
    }
    for (int i = middle; i < last; ++i) {
        if (key[i] > k[0]) {
            const float vk = key[i];
            const int vi = idx[i];
            key[i] = k[0]; idx[i] = x[0];
            adjust_heap_serial(k, x, 0, len, vk, vi);
        }
    }
    const float tk = key[first]; key[first] = key[nth]; key[nth] = tk;
    const int ti = idx[first]; idx[first] = idx[nth]; idx[nth] = ti;
}

// retainBest(keys, n_points) on key[0..n) (comparison: larger response first) with payload idx; returns the number kept.
// *heap_used (optional) is set when introselect's depth limit was reached and libstdc++'s heap select had to be replayed.
__device__ int retain_best_cta(float* key, int* idx, int* lpos, int* rpos, int n, int n_points, int* s, int* heap_used = nullptr) {
    if (n_points < 0 || n <= n_points) return n;
    if (n_points == 0) return 0;
    const int tid = threadIdx.x, nth = n_points - 1;
    int first = 0, last = n, depth = 2 * (31 - __clz(n));
    bool sorted_tail = true;
    while (last - first > 3) {
        if (depth == 0) {
            if (tid == 0) {
                heap_select_serial(key, idx, first, nth + 1, last, nth);
                if (heap_used) *heap_used = 1;
            }
            sorted_tail = false;                                 // std::__introselect returns here, without the insertion sort
            break;
        }
        --depth;
        if (tid == 0) {                                          // std::__move_median_to_first(first, first + 1, mid, last - 1), comp = greater
            const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
            const float ka = key[a], kb = key[b], kc = key[c];
            int m;
            if (ka > kb) m = kb > kc ? b : (ka > kc ? c : a);
            else m = ka > kc ? a : (kb > kc ? c : b);
            const float kf = key[first]; key[first] = key[m]; key[m] = kf;
            const int jf = idx[first]; idx[first] = idx[m]; idx[m] = jf;
        }
        __syncthreads();
        const float P = key[first];
        int tl, tr;
        // std::__unguarded_partition(first + 1, last, pivot): left scan passes elements with key > P, right scan those with P > key
        const int K = swap_pairs(key, idx, lpos, rpos, first + 1, last, [P](float v) { return !(v > P); }, [P](float v) { return !(P > v); }, s, tl, tr);
        int cut = 0x7fffffff;
        if (K < tl) cut = lpos[K];
        if (K >= 1) cut = min(cut, rpos[K - 1]);
        if (cut == 0x7fffffff) return -1;                        // cannot happen (median of three leaves a sentinel on either side)
        if (cut <= nth) first = cut; else last = cut;
        __syncthreads();
    }
    if (tid == 0 && sorted_tail) {                               // std::__insertion_sort(first, last), at most 3 elements
        for (int i = first + 1; i < last; ++i) {
            const float v = key[i];
            const int vi = idx[i];
            if (v > key[first]) {
                for (int j = i; j > first; --j) { key[j] = key[j - 1]; idx[j] = idx[j - 1]; }
                key[first] = v; idx[first] = vi;
            } else {
                int j = i;
                while (v > key[j - 1]) { key[j] = key[j - 1]; idx[j] = idx[j - 1]; --j; }
                key[j] = v; idx[j] = vi;
            }
        }
    }
    __syncthreads();
    // std::partition(begin + n_points, end, response >= ambiguous): ties with the n-th best response are kept as well
    const float amb = key[nth];
    int tl, tr;
    swap_pairs(key, idx, lpos, rpos, n_points, n, [amb](float v) { return !(v >= amb); }, [amb](float v) { return v >= amb; }, s, tl, tr);
    return n_points + tr;
}

// retainBest #1: per level, candidates by FAST score, 2 * nfeaturesPerLevel kept (+ ties)
__global__ void __launch_bounds__(kBorbThreads) k_borb_retain(const BorbPlan* __restrict__ plan, int stage, const float* __restrict__ resp_in,
                                                              const int* __restrict__ n_in, float* __restrict__ key, int* __restrict__ idx,
                                                              int* __restrict__ lpos, int* __restrict__ rpos, int* __restrict__ n_out,
                                                              int* __restrict__ err) {
    __shared__ int s[kBorbThreads / 32];
    const int l = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const BorbLevel& g = plan->lv[l];
    const size_t base = (size_t)b * plan->cand_total + g.cand_off;
    const int n = n_in[b * kBorbLevels + l];
    for (int i = tid; i < n; i += kBorbThreads) { key[base + i] = resp_in[base + i]; idx[base + i] = i; }
    __syncthreads();
    const int kept = retain_best_cta(key + base, idx + base, lpos + base, rpos + base, n, stage == 0 ? 2 * g.nfeat : g.nfeat, s);
    if (tid == 0) {
        n_out[b * kBorbLevels + l] = kept < 0 ? 0 : kept;
        if (kept < 0) atomicExch(err, 1);                        // partition found no cut: cannot happen, reported rather than trusted
    }
}

// stand-alone retainBest on caller data (parity tests of the replay against std::nth_element + std::partition)
__global__ void __launch_bounds__(kBorbThreads) k_borb_retain_debug(float* key, int* idx, int* lpos, int* rpos, int n, int n_points, int* n_out) {
    __shared__ int s[kBorbThreads / 32];
    for (int i = threadIdx.x; i < n; i += kBorbThreads) idx[i] = i;
    if (threadIdx.x == 0) n_out[1] = 0;
    __syncthreads();
    const int kept = retain_best_cta(key, idx, lpos, rpos, n, n_points, s, n_out + 1);
    if (threadIdx.x == 0) n_out[0] = kept;
}

// HarrisResponses of the keypoints kept by the first retainBest, in their kept order
__global__ void k_borb_harris(const BorbPlan* __restrict__ plan, const uint8_t* __restrict__ pyr, const uint32_t* __restrict__ cand_xy,
                              const int* __restrict__ idx1, const int* __restrict__ n1, float* __restrict__ resp2) {
    const int l = blockIdx.y, b = blockIdx.z, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1[b * kBorbLevels + l]) return;
    const BorbLevel& g = plan->lv[l];
    const size_t base = (size_t)b * plan->cand_total + g.cand_off;
    const uint32_t xy = cand_xy[base + idx1[base + i]];
    const int x0 = (int)(xy & 0xFFFFu), y0 = (int)(xy >> 16), step = g.pitch;
    const uint8_t* p0 = pyr + (size_t)b * plan->pyr_bytes + g.img_off + (size_t)(y0 - 3 + kBorbBorder) * step + (x0 - 3 + kBorbBorder);
    int a = 0, bb = 0, c = 0;
    for (int r = 0; r < 7; ++r)
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const uint8_t* p = p0 + r * step + q;
            const int Ix = (p[1] - p[-1]) * 2 + (p[-step + 1] - p[-step - 1]) + (p[step + 1] - p[step - 1]);
            const int Iy = (p[step] - p[-step]) * 2 + (p[step - 1] - p[-step - 1]) + (p[step + 1] - p[-step + 1]);
            a += Ix * Ix; bb += Iy * Iy; c += Ix * Iy;
        }
    const float scale = __fdiv_rn(1.f, __fmul_rn(28.f, 255.f));                  // 1.f / ((1 << 2) * blockSize * 255.f)
    const float ssss = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float fa = (float)a, fb = (float)bb, fc = (float)c;
    const float apb = __fadd_rn(fa, fb);
    // ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale_sq_sq
    const float r = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, apb), apb)), ssss);
    resp2[base + i] = r;
}

// second retainBest result -> ICAngles -> the keypoint records of detect(), levels concatenated.  One warp per keypoint: lane = row
// v = lane - 15 of the radius-15 disc (integer sums, so the order of addition is free).
__global__ void __launch_bounds__(128) k_borb_finish(const BorbPlan* __restrict__ plan, const uint8_t* __restrict__ pyr, const uint32_t* __restrict__ cand_xy,
                                                     const int* __restrict__ idx1, const int* __restrict__ idx2, const float* __restrict__ resp2,
                                                     const int* __restrict__ n2, int cap, fbe_keypoint* __restrict__ out, int* __restrict__ out_n,
                                                     int* __restrict__ err) {
    const int l = blockIdx.y, b = blockIdx.z, j = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int* cnt = n2 + b * kBorbLevels;
    int off = 0, total = 0;
    for (int i = 0; i < kBorbLevels; ++i) { if (i < l) off += cnt[i]; total += cnt[i]; }
    if (l == 0 && blockIdx.x == 0 && threadIdx.x == 0) { out_n[b] = min(total, cap); if (total > cap) atomicExch(err, 2); }
    if (j >= cnt[l] || off + j >= cap) return;
    const BorbLevel& g = plan->lv[l];
    const size_t base = (size_t)b * plan->cand_total + g.cand_off;
    const int i1 = idx2[base + j];
    const uint32_t xy = cand_xy[base + idx1[base + i1]];
    const int x0 = (int)(xy & 0xFFFFu), y0 = (int)(xy >> 16), step = g.pitch;
    const uint8_t* c = pyr + (size_t)b * plan->pyr_bytes + g.img_off + (size_t)(y0 + kBorbBorder) * step + (x0 + kBorbBorder);
    int m01 = 0, m10 = 0;
    if (lane < 31) {
        const int v = lane - 15, d = c_borb_umax[v < 0 ? -v : v];
        const uint8_t* row = c + v * step;
        int rs = 0;
        for (int u = -d; u <= d; ++u) { const int q = row[u]; rs += q; m10 += u * q; }
        m01 = v * rs;
    }
    m01 = __reduce_add_sync(0xffffffffu, m01);
    m10 = __reduce_add_sync(0xffffffffu, m10);
    if (lane == 0) {
        fbe_keypoint k;
        k.x = __fmul_rn((float)x0, g.scale); k.y = __fmul_rn((float)y0, g.scale);
        k.size = __fmul_rn(31.f, g.scale);
        k.angle = fast_atan2_deg((float)m01, (float)m10);
        k.response = resp2[base + i1];
        k.octave = l; k.class_id = -1;
        out[(size_t)b * cap + off + j] = k;
    }
}

// ---- compute() -------------------------------------------------------------------------------------------------------------------
// KeyPointsFilter::runByImageBorder(keypoints, image.size(), 31) on the ROUNDED positions, then (only when the input was not sorted
// by octave) a stable regrouping by octave.  One CTA per frame.  flags[b]: bit 0 = an octave outside the pyramid was present.
__global__ void __launch_bounds__(1024) k_borb_prefilter(const BorbPlan* __restrict__ plan, const fbe_keypoint* __restrict__ in,
                                                         const int* __restrict__ n_in, int cap, fbe_keypoint* __restrict__ out,
                                                         int* __restrict__ n_out, int* __restrict__ err) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n = min(n_in[b], cap);
    in += (size_t)b * cap; out += (size_t)b * cap;
    __shared__ int warp_sum[32];
    __shared__ int running, unsorted, bad;
    if (tid == 0) { running = 0; unsorted = 0; bad = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int o = in[i].octave;
        if (o < 0 || o >= kBorbLevels) bad = 1;
        if (i > 0 && o < in[i - 1].octave) unsorted = 1;
    }
    __syncthreads();
    if (bad) { if (tid == 0) { n_out[b] = 0; atomicExch(err, 3); } return; }
    const int npass = unsorted ? kBorbLevels : 1;
    for (int pass = 0; pass < npass; ++pass) {
        for (int base = 0; base < n; base += 1024) {
            const int i = base + tid;
            bool keep = false;
            fbe_keypoint k;
            if (i < n) {
                k = in[i];
                const int x = __float2int_rn(k.x), y = __float2int_rn(k.y);
                keep = x >= kBorbEdge && x < plan->cols - kBorbEdge && y >= kBorbEdge && y < plan->rows - kBorbEdge && (!unsorted || k.octave == pass);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) warp_sum[w] = __popc(bal);
            __syncthreads();
            int off = running;
            for (int q = 0; q < w; ++q) off += warp_sum[q];
            if (keep) out[off + __popc(bal & ((1u << lane) - 1u))] = k;
            __syncthreads();
            if (tid == 0) { int t = 0; for (int q = 0; q < 32; ++q) t += warp_sum[q]; running += t; }
            __syncthreads();
        }
    }
    if (tid == 0) n_out[b] = (plan->rows <= 2 * kBorbEdge || plan->cols <= 2 * kBorbEdge) ? 0 : running;
}

// float sepFilter2D Gaussian of a level ROI (see file header): tile 64 x 32, taps as float32 bit patterns of getGaussianKernel(7, 2, CV_32F)
__global__ void __launch_bounds__(256) k_borb_blur(const BorbPlan* __restrict__ plan, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur) {
    constexpr int TW = 64, TH = 32;
    __shared__ float hs[(TH + 6) * TW];
    int l = 0, t = blockIdx.x;                                   // tile index over all levels -> (level, tile)
    while (l < kBorbLevels - 1 && t >= plan->lv[l].blur_tiles) { t -= plan->lv[l].blur_tiles; ++l; }
    const BorbLevel& g = plan->lv[l];
    const int ntx = (g.w + TW - 1) / TW;
    const int b = blockIdx.y, x0 = (t % ntx) * TW, y0 = (t / ntx) * TH, tid = threadIdx.x;
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu), k2 = __uint_as_float(0x3e434a39u), k3 = __uint_as_float(0x3e5d4ae0u);
    const uint8_t* img = pyr + (size_t)b * plan->pyr_bytes + g.img_off + (size_t)kBorbBorder * g.pitch + kBorbBorder;
    for (int i = tid; i < (TH + 6) * TW; i += 256) {
        const int r = i / TW, cx = i - r * TW;
        const int x = min(x0 + cx, g.w - 1), y = min(y0 + r - 3, g.h + 2);      // clamped reads stay inside the 32-px frame; their results are unused
        const uint8_t* S = img + (size_t)y * g.pitch + (x - 3);
        float s = __fmul_rn(k0, (float)S[0]);
        s = fmaf(k1, (float)S[1], s); s = fmaf(k2, (float)S[2], s); s = fmaf(k3, (float)S[3], s);
        s = fmaf(k2, (float)S[4], s); s = fmaf(k1, (float)S[5], s); s = fmaf(k0, (float)S[6], s);
        hs[i] = s;
    }
    __syncthreads();
    for (int i = tid; i < TH * TW; i += 256) {
        const int r = i / TW, cx = i - r * TW;
        const int x = x0 + cx, y = y0 + r;
        if (x >= g.w || y >= g.h) continue;
        const float* c = hs + (r + 3) * TW + cx;
        float s = __fmul_rn(k3, c[0]);
        s = fmaf(k2, __fadd_rn(c[TW], c[-TW]), s);
        s = fmaf(k1, __fadd_rn(c[2 * TW], c[-2 * TW]), s);
        s = fmaf(k0, __fadd_rn(c[3 * TW], c[-3 * TW]), s);
        const int v = __float2int_rn(s);
        blur[(size_t)b * plan->pyr_bytes + g.img_off + (size_t)(y + kBorbBorder) * g.pitch + (x + kBorbBorder)] = (uint8_t)min(max(v, 0), 255);
    }
}

// steered BRIEF (computeOrbDescriptors, orb.cpp): one warp per keypoint, lane = output byte.  Samples inside the level ROI come
// from the blurred level, samples in the 32-px frame from the unblurred pyramid (cv::ORB blurs the ROI in place and leaves the frame).
__global__ void __launch_bounds__(128) k_borb_describe(const BorbPlan* __restrict__ plan, const uint8_t* __restrict__ pyr,
                                                       const uint8_t* __restrict__ blur, const fbe_keypoint* __restrict__ kps,
                                                       const int* __restrict__ n_arr, int cap, uint8_t* __restrict__ desc) {
    const int b = blockIdx.y, j = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= min(n_arr[b], cap)) return;
    const fbe_keypoint k = kps[(size_t)b * cap + j];
    const BorbLevel& g = plan->lv[k.octave];
    const float scale = __fdiv_rn(1.f, g.scale);
    const float ang = __fmul_rn(k.angle, (float)(3.14159265358979323846 / 180.0));
    float a = 0.f, s = 0.f;
    if (lane == 0) { a = (float)cos((double)ang); s = (float)sin((double)ang); }       // one fp64 evaluation per keypoint, not per lane
    a = __shfl_sync(0xffffffffu, a, 0); s = __shfl_sync(0xffffffffu, s, 0);
    const int cx = __float2int_rn(__fmul_rn(k.x, scale)), cy = __float2int_rn(__fmul_rn(k.y, scale));
    const size_t lvl = (size_t)b * plan->pyr_bytes + g.img_off;
    auto sample = [&](int px, int py) -> int {
        const float x = __fsub_rn(__fmul_rn((float)px, a), __fmul_rn((float)py, s));
        const float y = __fadd_rn(__fmul_rn((float)px, s), __fmul_rn((float)py, a));
        int ix = cx + __float2int_rn(x), iy = cy + __float2int_rn(y);
        const bool inside = ix >= 0 && ix < g.w && iy >= 0 && iy < g.h;
        ix = min(max(ix, -kBorbBorder), g.w + kBorbBorder - 1);                 // (defensive: keypoints handed to compute() cannot reach further)
        iy = min(max(iy, -kBorbBorder), g.h + kBorbBorder - 1);
        const size_t o = lvl + (size_t)(iy + kBorbBorder) * g.pitch + (ix + kBorbBorder);
        return inside ? blur[o] : pyr[o];
    };
    const int8_t* pat = c_borb_pattern + lane * 32;
    int val = 0;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const int t0 = sample(pat[4 * bit], pat[4 * bit + 1]);
        const int t1 = sample(pat[4 * bit + 2], pat[4 * bit + 3]);
        val |= (t0 < t1) << bit;
    }
    desc[((size_t)b * cap + j) * 32 + lane] = (uint8_t)val;
}

}  // namespace fbe

using namespace fbe;

// ---- host side -------------------------------------------------------------------------------------------------------------------
struct fbe_bird_orb {
    int device = 0, max_batch = 0, kp_cap = 0;
    BorbPlan hplan;
    BorbPlan* dplan = nullptr;
    short* d_band_level = nullptr;
    BorbTab* d_tabs = nullptr;
    cudaStream_t stream = nullptr;
    size_t fast_smem = 0;
    // per-frame slabs
    uint8_t *pyr = nullptr, *blur = nullptr, *mpyr = nullptr;
    uint8_t *d_img = nullptr, *d_mask = nullptr, *d_contour = nullptr;     // staging of the host images [B][rows][cols]
    uint2* band_slots = nullptr; int* band_count = nullptr;
    uint32_t* cand_xy = nullptr; float *cand_resp = nullptr, *resp2 = nullptr, *key = nullptr;
    int *idx1 = nullptr, *idx2 = nullptr, *lpos = nullptr, *rpos = nullptr, *cand_n = nullptr, *n1 = nullptr, *n2 = nullptr;
    fbe_keypoint *kps_a = nullptr, *kps_b = nullptr; int *n_a = nullptr, *n_b = nullptr, *n_c = nullptr;
    uint8_t* desc = nullptr;
    uint8_t* keep = nullptr; int* iters = nullptr; float* subpix_mask = nullptr;
    int* err = nullptr;
    bool pyr_valid = false;            // the pyramid on the device belongs to the images of the last detect / upload
};

namespace {

#define FBE_TRY(expr) do { int _rc = (expr); if (_rc != FBE_OK) return _rc; } while (0)

void exact_tab(int src, int dst, std::vector<BorbTab>& out) {
    // interpolationLinear<uchar>::getCoeffs of cv::resize INTER_LINEAR_EXACT (softdouble == IEEE double arithmetic)
    const double scale = 1.0 / ((double)dst / src);
    for (int v = 0; v < dst; ++v) {
        const double fval = scale * ((double)v + 0.5) - 0.5;
        const int ival = (int)std::floor(fval);
        BorbTab t;
        if (ival >= 0 && src > 1) {
            if (ival < src - 1) {
                const int c1 = (int)std::nearbyint((fval - (double)ival) * 256.0);
                t.o0 = (short)ival; t.o1 = (short)(ival + 1); t.c1 = (short)c1; t.c0 = (short)(256 - c1);
            } else { t.o0 = t.o1 = (short)(src - 1); t.c0 = 256; t.c1 = 0; }
        } else { t.o0 = t.o1 = 0; t.c0 = 256; t.c1 = 0; }
        out.push_back(t);
    }
}

int build_borb_plan(int nfeatures, int rows, int cols, BorbPlan& p, std::vector<short>& band_level, std::vector<BorbTab>& tabs) {
    std::memset(&p, 0, sizeof(p));
    p.rows = rows; p.cols = cols; p.nfeatures = nfeatures;
    if (rows < 1 || cols < 1 || rows > 4095 || cols > 4095) { set_error("bird image size out of range (1 .. 4095 per side)"); return FBE_E_UNSUPPORTED; }
    // nfeaturesPerLevel (computeKeyPoints, orb.cpp)
    const float factor = (float)(1.0 / (double)1.2f);
    float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)kBorbLevels));
    int sum = 0;
    for (int l = 0; l < kBorbLevels - 1; ++l) {
        p.lv[l].nfeat = (int)std::nearbyint(nd);
        sum += p.lv[l].nfeat;
        nd *= factor;
    }
    p.lv[kBorbLevels - 1].nfeat = std::max(nfeatures - sum, 0);
    int img_off = 0, mask_off = 0, band_off = 0, cand_off = 0;
    for (int l = 0; l < kBorbLevels; ++l) {
        BorbLevel& g = p.lv[l];
        g.scale = (float)std::pow((double)1.2f, (double)l);
        const float inv = 1.0f / g.scale;
        g.w = (int)std::nearbyintf(cols * inv); g.h = (int)std::nearbyintf(rows * inv);
        if (g.w < 1 || g.h < 1) { set_error("bird pyramid level collapses to zero size"); return FBE_E_UNSUPPORTED; }
        g.pitch = (g.w + 2 * kBorbBorder + 15) & ~15;
        g.ph = g.h + 2 * kBorbBorder;
        g.img_off = img_off; img_off += (g.pitch * g.ph + 255) & ~255;
        g.mask_off = mask_off; mask_off += (g.w * g.h + 255) & ~255;
        g.kw = g.w - 2 * kBorbEdge; g.kh = g.h - 2 * kBorbEdge;
        g.nbands = (g.kw > 0 && g.kh > 0) ? (g.kh + kBorbBand - 1) / kBorbBand : 0;
        if (g.nbands > 512) { set_error("bird image too tall"); return FBE_E_UNSUPPORTED; }
        g.band_off = band_off; band_off += g.nbands;
        g.band_cap = g.nbands ? kBorbBand * ((g.kw + 1) / 2) : 0;              // NMS survivors are never horizontal neighbours
        g.cand_cap = g.nbands ? ((g.kw + 1) / 2) * ((g.kh + 1) / 2) + 8 : 8;  // ... nor vertical ones
        g.cand_off = cand_off; cand_off += g.cand_cap;
        for (int i = 0; i < g.nbands; ++i) band_level.push_back((short)l);
        g.blur_tiles = ((g.w + 63) / 64) * ((g.h + 31) / 32);
        if (l > 0) {
            g.tabx_off = (int)tabs.size(); exact_tab(p.lv[l - 1].w, g.w, tabs);
            g.taby_off = (int)tabs.size(); exact_tab(p.lv[l - 1].h, g.h, tabs);
        }
    }
    p.pyr_bytes = img_off; p.mask_bytes = mask_off; p.bands_total = band_off; p.cand_total = cand_off;
    p.band_slots_total = band_off * p.lv[0].band_cap;
    return FBE_OK;
}

template <class T> int dalloc(T** p, size_t count) {
    FBE_CUDA(cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T)));
    return FBE_OK;
}

int upload_images(fbe_bird_orb* h, uint8_t* dst, const uint8_t* src, size_t step, size_t stride, int B) {
    const int rows = h->hplan.rows, cols = h->hplan.cols;
    if (stride == step * (size_t)rows || B == 1) {
        FBE_CUDA(cudaMemcpy2DAsync(dst, (size_t)cols, src, step, (size_t)cols, (size_t)rows * B, cudaMemcpyHostToDevice, h->stream));
    } else {
        for (int b = 0; b < B; ++b)
            FBE_CUDA(cudaMemcpy2DAsync(dst + (size_t)rows * cols * b, (size_t)cols, src + stride * b, step, (size_t)cols, (size_t)rows, cudaMemcpyHostToDevice, h->stream));
    }
    return FBE_OK;
}

// image (+ mask) pyramid of B frames from the staged device images
int run_pyramid(fbe_bird_orb* h, int B, bool has_mask) {
    const BorbPlan& p = h->hplan;
    cudaStream_t st = h->stream;
    {
        const BorbLevel& g = p.lv[0];
        dim3 grid((g.w + 2 * kBorbBorder + 127) / 128, g.ph, B);
        k_borb_level0<<<grid, 128, 0, st>>>(h->dplan, h->d_img, (size_t)p.cols, (size_t)p.rows * p.cols, has_mask ? h->d_mask : nullptr, (size_t)p.cols,
                                            (size_t)p.rows * p.cols, h->pyr, h->mpyr);
        count_launch();
    }
    for (int l = 1; l < kBorbLevels; ++l) {
        const BorbLevel& g = p.lv[l];
        k_borb_resize<<<dim3((g.w + 2 * kBorbBorder + 127) / 128, g.ph + (has_mask ? g.h : 0), B), 128, 0, st>>>(h->dplan, h->d_tabs, l, h->pyr, h->mpyr);
        count_launch();
    }
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

// detect() on the device pyramid -> h->kps_a / h->n_a
int run_detect(fbe_bird_orb* h, int B, bool has_mask) {
    const BorbPlan& p = h->hplan;
    cudaStream_t st = h->stream;
    if (p.bands_total == 0) { FBE_CUDA(cudaMemsetAsync(h->n_a, 0, (size_t)B * sizeof(int), st)); return FBE_OK; }
    k_borb_fast<<<dim3(p.bands_total, B), kBorbThreads, h->fast_smem, st>>>(h->dplan, h->d_band_level, h->pyr, h->mpyr, has_mask, h->band_slots, h->band_count);
    k_borb_gather<<<dim3(kBorbLevels, B), kBorbThreads, 0, st>>>(h->dplan, h->band_slots, h->band_count, h->cand_xy, h->cand_resp, h->cand_n);
    k_borb_retain<<<dim3(kBorbLevels, B), kBorbThreads, 0, st>>>(h->dplan, 0, h->cand_resp, h->cand_n, h->key, h->idx1, h->lpos, h->rpos, h->n1, h->err);
    int max_cap = 0;
    for (int l = 0; l < kBorbLevels; ++l) max_cap = std::max(max_cap, p.lv[l].cand_cap);
    const int max_kp = std::min(max_cap, h->kp_cap);                  // a level can not contribute more than the output holds
    k_borb_harris<<<dim3((max_cap + 127) / 128, kBorbLevels, B), 128, 0, st>>>(h->dplan, h->pyr, h->cand_xy, h->idx1, h->n1, h->resp2);
    k_borb_retain<<<dim3(kBorbLevels, B), kBorbThreads, 0, st>>>(h->dplan, 1, h->resp2, h->n1, h->key, h->idx2, h->lpos, h->rpos, h->n2, h->err);
    k_borb_finish<<<dim3((max_kp + 3) / 4, kBorbLevels, B), 128, 0, st>>>(h->dplan, h->pyr, h->cand_xy, h->idx1, h->idx2, h->resp2, h->n2, h->kp_cap,
                                                                            h->kps_a, h->n_a, h->err);
    count_launch(6);
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

// compute() on the device pyramid: d_in / d_nin -> h->kps_b / h->n_b (filtered, regrouped) + h->desc
int run_compute(fbe_bird_orb* h, int B, const fbe_keypoint* d_in, const int* d_nin) {
    const BorbPlan& p = h->hplan;
    cudaStream_t st = h->stream;
    k_borb_prefilter<<<B, 1024, 0, st>>>(h->dplan, d_in, d_nin, h->kp_cap, h->kps_b, h->n_b, h->err);
    count_launch();
    int tiles = 0;
    for (int l = 0; l < kBorbLevels; ++l) tiles += p.lv[l].blur_tiles;
    k_borb_blur<<<dim3(tiles, B), 256, 0, st>>>(h->dplan, h->pyr, h->blur);
    count_launch();
    k_borb_describe<<<dim3((h->kp_cap + 3) / 4, B), 128, 0, st>>>(h->dplan, h->pyr, h->blur, h->kps_b, h->n_b, h->kp_cap, h->desc);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int check_err(fbe_bird_orb* h) {
    int e = 0;
    FBE_CUDA(cudaMemcpyAsync(&e, h->err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FBE_CUDA(cudaStreamSynchronize(h->stream));
    if (e) FBE_CUDA(cudaMemsetAsync(h->err, 0, sizeof(int), h->stream));
    if (e == 1) { set_error("retainBest: a partition pass of the introselect replay found no cut (internal error)"); return FBE_E_UNSUPPORTED; }
    if (e == 2) { set_error("more keypoints than the output capacity (ties of the n-th best response are all kept)"); return FBE_E_CAPACITY; }
    if (e == 3) { set_error("keypoint octave outside 0 .. 7"); return FBE_E_INVALID; }
    return FBE_OK;
}

// results to the host: the counts (and the error word) first, then only the used front of every fixed-stride list
int download(fbe_bird_orb* h, int B, const int* d_n, const fbe_keypoint* d_kps, const uint8_t* d_desc, int32_t* n, fbe_keypoint* kps, uint8_t* desc) {
    FBE_CUDA(cudaMemcpyAsync(n, d_n, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    FBE_TRY(check_err(h));                                       // synchronises the stream
    int nmax = 0;
    for (int b = 0; b < B; ++b) nmax = std::max(nmax, (int)n[b]);
    if (nmax == 0) return FBE_OK;
    const size_t cap = (size_t)h->kp_cap;
    FBE_CUDA(cudaMemcpy2DAsync(kps, cap * sizeof(fbe_keypoint), d_kps, cap * sizeof(fbe_keypoint), (size_t)nmax * sizeof(fbe_keypoint), (size_t)B,
                               cudaMemcpyDeviceToHost, h->stream));
    if (desc) FBE_CUDA(cudaMemcpy2DAsync(desc, cap * 32, d_desc, cap * 32, (size_t)nmax * 32, (size_t)B, cudaMemcpyDeviceToHost, h->stream));
    FBE_CUDA(cudaStreamSynchronize(h->stream));
    return FBE_OK;
}

}  // namespace

extern "C" {

int fbe_bird_orb_create(int32_t nfeatures, int32_t rows, int32_t cols, int32_t max_batch, int32_t device, fbe_bird_orb** out) {
    if (!out || nfeatures < 1 || max_batch < 1) return FBE_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device: this library has no CPU path"); return FBE_E_CUDA; }
    FBE_CUDA(cudaSetDevice(device));
    fbe_bird_orb* h = new (std::nothrow) fbe_bird_orb();
    if (!h) return FBE_E_INVALID;
    h->device = device; h->max_batch = max_batch;
    std::vector<short> band_level;
    std::vector<BorbTab> tabs;
    int rc = build_borb_plan(nfeatures, rows, cols, h->hplan, band_level, tabs);
    if (rc != FBE_OK) { delete h; return rc; }
    const BorbPlan& p = h->hplan;
    h->kp_cap = std::max(4 * nfeatures, 4096);
    const size_t B = (size_t)max_batch, ct = (size_t)p.cand_total, cap = (size_t)h->kp_cap, img = (size_t)rows * cols;
    auto fail = [&](int code) { fbe_bird_orb_destroy(h); return code; };
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); return fail(FBE_E_CUDA); }
    int ok = FBE_OK;
#define A(ptr, count) if (ok == FBE_OK) ok = dalloc(&h->ptr, (count))
    A(dplan, 1); A(d_band_level, band_level.size()); A(d_tabs, tabs.size());
    A(pyr, B * p.pyr_bytes); A(blur, B * p.pyr_bytes); A(mpyr, B * p.mask_bytes);
    A(d_img, B * img); A(d_mask, B * img); A(d_contour, B * img);
    A(band_slots, B * p.band_slots_total); A(band_count, B * p.bands_total);
    A(cand_xy, B * ct); A(cand_resp, B * ct); A(resp2, B * ct); A(key, B * ct);
    A(idx1, B * ct); A(idx2, B * ct); A(lpos, B * ct); A(rpos, B * ct);
    A(cand_n, B * kBorbLevels); A(n1, B * kBorbLevels); A(n2, B * kBorbLevels);
    A(kps_a, B * cap); A(kps_b, B * cap); A(n_a, B); A(n_b, B); A(n_c, B);
    A(desc, B * cap * 32); A(keep, B * cap); A(iters, B * cap); A(subpix_mask, 441); A(err, 1);
#undef A
    if (ok != FBE_OK) return fail(ok);
    if (cudaMemcpy(h->dplan, &h->hplan, sizeof(BorbPlan), cudaMemcpyHostToDevice) != cudaSuccess ||
        (!band_level.empty() && cudaMemcpy(h->d_band_level, band_level.data(), band_level.size() * sizeof(short), cudaMemcpyHostToDevice) != cudaSuccess) ||
        (!tabs.empty() && cudaMemcpy(h->d_tabs, tabs.data(), tabs.size() * sizeof(BorbTab), cudaMemcpyHostToDevice) != cudaSuccess) ||
        cudaMemset(h->err, 0, sizeof(int)) != cudaSuccess || cudaMemset(h->blur, 0, B * p.pyr_bytes) != cudaSuccess) {
        set_error("bird ORB plan upload failed");
        return fail(FBE_E_CUDA);
    }
    const int sw = p.lv[0].kw + 2, words = (std::max(p.lv[0].kw, 1) + 31) >> 5;
    h->fast_smem = (size_t)(((kBorbBand + 2) * std::max(sw, 1) + 15) & ~15) + (size_t)kBorbBand * words * 4;
    if (h->fast_smem > 200 * 1024) { set_error("bird image too wide for the FAST band kernel"); return fail(FBE_E_UNSUPPORTED); }
    if (h->fast_smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_borb_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fast_smem) != cudaSuccess) {
        set_error("shared memory configuration failed");
        return fail(FBE_E_CUDA);
    }
    *out = h;
    return FBE_OK;
}

int fbe_bird_orb_destroy(fbe_bird_orb* h) {
    if (!h) return FBE_E_INVALID;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void* ptrs[] = {h->dplan, h->d_band_level, h->d_tabs, h->pyr, h->blur, h->mpyr, h->d_img, h->d_mask, h->d_contour, h->band_slots, h->band_count,
                    h->cand_xy, h->cand_resp, h->resp2, h->key, h->idx1, h->idx2, h->lpos, h->rpos, h->cand_n, h->n1, h->n2, h->kps_a, h->kps_b,
                    h->n_a, h->n_b, h->n_c, h->desc, h->keep, h->iters, h->subpix_mask, h->err};
    for (void* q : ptrs) if (q) cudaFree(q);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return FBE_OK;
}

int fbe_bird_orb_max_keypoints(const fbe_bird_orb* h, int32_t* cap) {
    if (!h || !cap) return FBE_E_INVALID;
    *cap = h->kp_cap;
    return FBE_OK;
}

int fbe_debug_retain_best(const float* response, int32_t n, int32_t n_points, int32_t device, int32_t* order, int32_t* n_kept,
                          int32_t* heap_select_used) {
    if (!response || !order || !n_kept || n < 0) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(device));
    float* key = nullptr; int *idx = nullptr, *lp = nullptr, *rp = nullptr, *no = nullptr;
    const size_t nn = (size_t)std::max(n, 1);
    int rc = FBE_OK;
    if (cudaMalloc(&key, nn * 4) != cudaSuccess || cudaMalloc(&idx, nn * 4) != cudaSuccess || cudaMalloc(&lp, nn * 4) != cudaSuccess ||
        cudaMalloc(&rp, nn * 4) != cudaSuccess || cudaMalloc(&no, 8) != cudaSuccess) { set_error("allocation failed"); rc = FBE_E_CUDA; }
    if (rc == FBE_OK) {
        cudaMemcpy(key, response, (size_t)n * 4, cudaMemcpyHostToDevice);
        k_borb_retain_debug<<<1, kBorbThreads>>>(key, idx, lp, rp, n, n_points, no);
        count_launch();
        int h_no[2] = {0, 0};
        cudaMemcpy(h_no, no, 8, cudaMemcpyDeviceToHost);
        *n_kept = h_no[0];
        if (heap_select_used) *heap_select_used = h_no[1];
        if (cudaMemcpy(order, idx, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess || cudaGetLastError() != cudaSuccess) { set_error("retain_best debug run failed"); rc = FBE_E_CUDA; }
    }
    cudaFree(key); cudaFree(idx); cudaFree(lp); cudaFree(rp); cudaFree(no);
    return rc;
}

int fbe_bird_orb_detect(fbe_bird_orb* h, const uint8_t* imgs, size_t step, size_t stride, const uint8_t* masks, size_t mask_step,
                        size_t mask_stride, int32_t nframes, fbe_keypoint* kps, int32_t* n) {
    if (!h || !imgs || !kps || !n || nframes < 1 || nframes > h->max_batch || step < (size_t)h->hplan.cols || (masks && mask_step < (size_t)h->hplan.cols)) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(h->device));
    FBE_TRY(upload_images(h, h->d_img, imgs, step, stride, nframes));
    if (masks) FBE_TRY(upload_images(h, h->d_mask, masks, mask_step, mask_stride, nframes));
    FBE_TRY(run_pyramid(h, nframes, masks != nullptr));
    FBE_TRY(run_detect(h, nframes, masks != nullptr));
    return download(h, nframes, h->n_a, h->kps_a, nullptr, n, kps, nullptr);
}

int fbe_bird_orb_compute(fbe_bird_orb* h, const uint8_t* imgs, size_t step, size_t stride, int32_t nframes, fbe_keypoint* kps, int32_t* n,
                         uint8_t* desc) {
    if (!h || !imgs || !kps || !n || !desc || nframes < 1 || nframes > h->max_batch || step < (size_t)h->hplan.cols) return FBE_E_INVALID;
    for (int b = 0; b < nframes; ++b) if (n[b] < 0 || n[b] > h->kp_cap) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(h->device));
    FBE_TRY(upload_images(h, h->d_img, imgs, step, stride, nframes));
    FBE_TRY(run_pyramid(h, nframes, false));
    FBE_CUDA(cudaMemcpyAsync(h->kps_a, kps, (size_t)nframes * h->kp_cap * sizeof(fbe_keypoint), cudaMemcpyHostToDevice, h->stream));
    FBE_CUDA(cudaMemcpyAsync(h->n_a, n, (size_t)nframes * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    FBE_TRY(run_compute(h, nframes, h->kps_a, h->n_a));
    return download(h, nframes, h->n_b, h->kps_b, h->desc, n, kps, desc);
}

int fbe_bird_features(fbe_bird_orb* h, const uint8_t* imgs, size_t step, size_t stride, const uint8_t* masks, size_t mask_step, size_t mask_stride,
                      const uint8_t* contours, size_t contour_step, size_t contour_stride, int32_t nframes, fbe_keypoint* kps, int32_t* n,
                      uint8_t* desc, int32_t* n_detected) {
    if (!h || !imgs || !kps || !n || !desc || nframes < 1 || nframes > h->max_batch || step < (size_t)h->hplan.cols ||
        (masks && mask_step < (size_t)h->hplan.cols) || (contours && contour_step < (size_t)h->hplan.cols)) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(h->device));
    const BorbPlan& p = h->hplan;
    FBE_TRY(upload_images(h, h->d_img, imgs, step, stride, nframes));
    if (masks) FBE_TRY(upload_images(h, h->d_mask, masks, mask_step, mask_stride, nframes));
    if (contours) FBE_TRY(upload_images(h, h->d_contour, contours, contour_step, contour_stride, nframes));
    FBE_TRY(run_pyramid(h, nframes, masks != nullptr));
    FBE_TRY(run_detect(h, nframes, masks != nullptr));                    // extractorBird->detect(mBirdviewImg, preKeysBird, mBirdviewMask)
    if (n_detected) FBE_CUDA(cudaMemcpyAsync(n_detected, h->n_a, (size_t)nframes * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    fbe_keypoint* d_res = nullptr;
    const int* d_res_n = nullptr;                                          // GuidenceKeyBirdPts + cv::cornerSubPix(.., Size(5,5), Size(-1,-1), {40, 0.001})
    FBE_TRY(launch_bird_refine_dev(contours ? h->d_contour : nullptr, h->d_img, p.rows, p.cols, nframes, h->kps_a, h->n_a, h->kp_cap, 5, 5, 40, 0.001,
                                   h->keep, h->kps_b, h->n_c, h->iters, h->subpix_mask, &d_res, &d_res_n, h->stream));
    if (d_res == h->kps_b) {      // run_compute writes kps_b: move the refined list out of its way
        FBE_CUDA(cudaMemcpyAsync(h->kps_a, h->kps_b, (size_t)nframes * h->kp_cap * sizeof(fbe_keypoint), cudaMemcpyDeviceToDevice, h->stream));
        d_res = h->kps_a;
    }
    FBE_TRY(run_compute(h, nframes, d_res, d_res_n));                     // extractorBird->compute(mBirdviewImg, mvKeysBird, mDescriptorsBird)
    return download(h, nframes, h->n_b, h->kps_b, h->desc, n, kps, desc);
}

}  // extern "C"
