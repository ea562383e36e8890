// 256-bit Hamming searches over grid-windowed candidates (ORBmatcher family) for sm_100a.
//
// Shape shared by every search: one WARP per query walks the grid window of
// Frame::GetFeaturesInArea[Birdview] (src/Frame.cc:493-546, 572-626) -- ix outer, iy inner, in-cell insertion order.
// Because the CSR grid is stored [ix*grows + iy], the cells (ix, minY..maxY) of one column are ONE contiguous item
// range, so the walk is a handful of coalesced range scans, lanes = candidates.  Candidates are ranked in traversal
// order with ballot/popc so that strict-'<' tie-breaking of the reference is reproduced: best = lexicographic
// min of (distance, rank), second = second element of that order.
// Order-INDEPENDENT searches (BirdviewMatch, BirdMapPointMatch, brute force) reduce to top-2 in registers.
// Order-DEPENDENT searches (SearchForInitialization, SearchByProjection x2, SearchByBoW) first store every
// (candidate, distance) row in traversal order -- all the Hamming work, fully parallel -- and then ONE warp per
// problem replays the reference's sequential gate / steal / assign rules over those rows.
#include <algorithm>
#include <climits>
#include "match_kernels.cuh"

namespace fbe {

__device__ __forceinline__ int hamming256(const uint32_t (&a)[8], const uint8_t* __restrict__ b) {
    const uint4 b0 = *reinterpret_cast<const uint4*>(b);
    const uint4 b1 = *reinterpret_cast<const uint4*>(b + 16);
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

__device__ __forceinline__ void load_desc(uint32_t (&a)[8], const uint8_t* __restrict__ p) {
    const uint4 v0 = *reinterpret_cast<const uint4*>(p);
    const uint4 v1 = *reinterpret_cast<const uint4*>(p + 16);
    a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w; a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
}

// keep the two smallest packed keys
__device__ __forceinline__ void top2_push(unsigned& k1, unsigned& k2, unsigned k) {
    if (k < k1) { k2 = k1; k1 = k; }
    else if (k < k2) k2 = k;
}
// two smallest keys of the warp, in every lane: two REDUX.MIN instead of five shuffle rounds.  Keys are distinct (the
// rank field is unique) except for kNoKey, so the runner-up is the smallest key that is not the winner.
__device__ __forceinline__ void top2_warp_merge(unsigned& k1, unsigned& k2) {
    const unsigned w = __reduce_min_sync(0xffffffffu, k1);
    const unsigned r = __reduce_min_sync(0xffffffffu, k1 == w ? k2 : k1);
    k1 = w;
    k2 = r;
}

// Warp-collective walk of the window.  fn(idx, pass, rank) is called for every lane of every chunk; `pass` lanes carry
// a candidate index that cleared the level and window filters, `rank` its position in traversal order.
template <class Fn>
__device__ __forceinline__ int window_walk(const fbe_keypoint* __restrict__ kps, const int* __restrict__ start,
                                           const int* __restrict__ items, float min_x, float min_y, float inv_w, float inv_h,
                                           int gcols, int grows, float x, float y, float r, int minL, int maxL,
                                           bool incl, Fn&& fn) {
    const int lane = threadIdx.x & 31;
    const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, min_x), r), inv_w)));
    if (cx0 >= gcols) return 0;
    const int cx1 = min(gcols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, min_x), r), inv_w)));
    if (cx1 < 0) return 0;
    const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, min_y), r), inv_h)));
    if (cy0 >= grows) return 0;
    const int cy1 = min(grows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, min_y), r), inv_h)));
    if (cy1 < 0) return 0;
    const bool check = (minL > 0) || (maxL >= 0);
    const int xe = incl ? cx1 + 1 : cx1, ye = incl ? cy1 + 1 : cy1;     // exclusive ends (bird: quirk Q2)
    int rank = 0;
    if (ye <= cy0) return 0;
    // The window's columns are FLATTENED into one candidate sequence: lane j fetches the item range of column cx0 + j (one
    // load latency for all columns instead of one per column), a warp scan turns the range lengths into offsets, and the
    // candidates are then taken 32 at a time across column boundaries -- traversal order (ix outer, iy inner, in-cell order)
    // is the flattened order, and no lane idles because one column happens to hold five keypoints.
    for (int cbase = cx0; cbase < xe; cbase += 32) {                 // 32 columns per pass (a window is rarely wider)
        const int ix = cbase + lane;
        int beg = 0, cnt = 0;
        if (ix < xe) {
            beg = start[ix * grows + cy0];
            cnt = start[ix * grows + ye] - beg;
        }
        int inc = cnt;                                              // inclusive prefix of the range lengths
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        const int exc = inc - cnt;
        for (int pbase = 0; pbase < total; pbase += 32) {
            const int p = pbase + lane;
            // column of candidate p = number of columns whose inclusive prefix is <= p (upper bound by bisection over the lanes)
            int col = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, inc, col + step - 1);
                if (v <= p) col += step;
            }
            col = min(col, 31);
            const int cbeg = __shfl_sync(0xffffffffu, beg, col), cexc = __shfl_sync(0xffffffffu, exc, col);
            int idx = -1;
            bool pass = false;
            if (p < total) {
                idx = items[cbeg + (p - cexc)];
                const fbe_keypoint* kp = kps + idx;
                const int oct = kp->octave;
                bool ok = true;
                if (check) {
                    if (oct < minL) ok = false;
                    if (maxL >= 0 && oct > maxL) ok = false;
                }
                if (ok) {
                    const float dx = __fsub_rn(kp->x, x), dy = __fsub_rn(kp->y, y);
                    pass = fabsf(dx) < r && fabsf(dy) < r;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            fn(idx, pass, rank + __popc(bal & ((1u << lane) - 1u)));
            rank += __popc(bal);
        }
    }
    return rank;
}

// ---- query construction ------------------------------------------------------------------------------------------
// SearchForInitialization / BirdviewMatch(isProject=0): one query per octave-0 keypoint of the reference frame,
// centred on vbPrevMatched[i] (pos != NULL) or on the keypoint itself; levels (0,0); others are marked r < 0.
__global__ void k_queries_from_kps(const fbe_keypoint* __restrict__ kps, const float2* __restrict__ pos, const int* __restrict__ n,
                                   int stride, float window, float4* __restrict__ q, int2* __restrict__ lv) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n[b]) return;
    const fbe_keypoint kp = kps[(size_t)b * stride + i];
    float x = kp.x, y = kp.y;
    if (pos) { const float2 p = pos[(size_t)b * stride + i]; x = p.x; y = p.y; }
    q[(size_t)b * stride + i] = make_float4(x, y, kp.octave > 0 ? -1.f : window, 0.f);
    lv[(size_t)b * stride + i] = make_int2(kp.octave, kp.octave);
}

// ---- parallel stage: rows of (candidate, distance) in traversal order -----------------------------------------------
__global__ void __launch_bounds__(256, 6) k_window_rows(FrameDev f, QueryDev qs, bool incl, int C, unsigned* __restrict__ rows,
                                                     int* __restrict__ cnt, int* __restrict__ overflow) {
    const int b = blockIdx.y;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= qs.nq[b]) return;
    const size_t qoff = (size_t)b * qs.stride + qi;
    const float4 q = qs.q[qoff];
    int total = 0;
    if (q.z >= 0.f) {
        const int2 lv = qs.lv[qoff];
        uint32_t qd[8];
        load_desc(qd, qs.desc + qoff * 32);
        const uint8_t* tdesc = f.desc + (size_t)b * f.kp_stride * 32;
        unsigned* row = rows + qoff * C;
        total = window_walk(f.kps + (size_t)b * f.kp_stride, f.start + (size_t)b * (f.gcols * f.grows + 1),
                            f.items + (size_t)b * f.kp_stride, f.min_x, f.min_y, f.inv_w, f.inv_h, f.gcols, f.grows, q.x, q.y,
                            q.z, lv.x, lv.y, incl, [&](int idx, bool pass, int rank) {
                                if (pass && rank < C) {
                                    const int d = hamming256(qd, tdesc + (size_t)idx * 32);
                                    row[rank] = ((unsigned)idx << kRowDistBits) | (unsigned)d;
                                }
                            });
    }
    if ((threadIdx.x & 31) == 0) {
        cnt[qoff] = min(total, C);
        if (total > C) {                      // overflow[0] = flag, overflow[4] = 1 + (problem << 16 | query) of one offender
            atomicExch(overflow, 1);
            atomicMax(overflow + 4, ((b << 16) | min(qi, 0xFFFF)) + 1);
        }
    }
}

// ---- parallel stage, order-independent searches: top-2 directly ------------------------------------------------------
// kReproj adds the per-candidate reprojection gate of ORBmatcher::Fuse (src/ORBmatcher.cc:911-937): q.w carries ur, the
// chi-square bounds are compared in double like the reference's `e2*invSigma2 > 5.99` (float product, double literal).
template <bool kReproj>
__global__ void __launch_bounds__(256, 6) k_window_top2(FrameDev f, QueryDev qs, bool incl, ReprojGate rg, int* __restrict__ best_idx,
                                                     int* __restrict__ best_dist, int* __restrict__ second_dist) {
    const int b = blockIdx.y;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= qs.nq[b]) return;
    const size_t qoff = (size_t)b * qs.stride + qi;
    const float4 q = qs.q[qoff];
    unsigned k1 = kNoKey, k2 = kNoKey;
    int my_idx = -1;       // candidate index behind this lane's k1
    if (q.z >= 0.f) {
        const int2 lv = qs.lv[qoff];
        uint32_t qd[8];
        load_desc(qd, qs.desc + qoff * 32);
        const uint8_t* tdesc = f.desc + (size_t)b * f.kp_stride * 32;
        window_walk(f.kps + (size_t)b * f.kp_stride, f.start + (size_t)b * (f.gcols * f.grows + 1),
                    f.items + (size_t)b * f.kp_stride, f.min_x, f.min_y, f.inv_w, f.inv_h, f.gcols, f.grows, q.x, q.y, q.z, lv.x,
                    lv.y, incl, [&](int idx, bool pass, int rank) {
                        if (kReproj && pass) {
                            const fbe_keypoint* kp = f.kps + (size_t)b * f.kp_stride + idx;
                            const float ex = __fsub_rn(q.x, kp->x), ey = __fsub_rn(q.y, kp->y);
                            float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                            const float kpr = rg.t_uright ? rg.t_uright[idx] : -1.0f;
                            double bound = 5.99;
                            if (kpr >= 0.0f) {
                                const float er = __fsub_rn(q.w, kpr);
                                e2 = __fadd_rn(e2, __fmul_rn(er, er));
                                bound = 7.8;
                            }
                            if ((double)__fmul_rn(e2, rg.inv_sigma2[kp->octave]) > bound) pass = false;
                        }
                        if (pass) {
                            const unsigned k = ((unsigned)hamming256(qd, tdesc + (size_t)idx * 32) << 20) | (unsigned)min(rank, 0xFFFFF);
                            if (k < k1) my_idx = idx;
                            top2_push(k1, k2, k);
                        }
                    });
    }
    const unsigned mine = k1;
    top2_warp_merge(k1, k2);
    // the lane that owns the winning key publishes its index
    const unsigned owner = __ballot_sync(0xffffffffu, mine == k1 && k1 != kNoKey);
    const int widx = __shfl_sync(0xffffffffu, my_idx, owner ? (__ffs(owner) - 1) : 0);
    if ((threadIdx.x & 31) == 0) {
        best_idx[qoff] = k1 == kNoKey ? -1 : widx;
        best_dist[qoff] = k1 == kNoKey ? INT_MAX : (int)(k1 >> 20);
        second_dist[qoff] = k2 == kNoKey ? INT_MAX : (int)(k2 >> 20);
    }
}

// ---- rotation histogram helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ int rot_bin(float a1, float a2) {
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / FBE_HISTO_LENGTH));
    if (bin == FBE_HISTO_LENGTH) bin = 0;
    return bin;
}

// ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1905-1946)
__device__ inline void three_maxima(const int* histo, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < FBE_HISTO_LENGTH; ++i) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// ---- sequential stage: one CTA per problem replays the reference's loop over the stored rows ----------------------
// The reference's loop is serial: each accepted match changes which candidates later queries may take (vMatchedDistance
// for SearchForInitialization, the taken flags elsewhere).  But an accept touches ONE target, and it changes the outcome
// of a later query only when that target is the later query's best or second-best candidate -- rare.  So the replay runs
// in waves of kResolveWave queries: every warp evaluates one query of the wave against the state left by the previous
// wave (top-2 over its row), then warp 0 takes one lane per query, decides accept / reject for all of them at once, and
// finds the first query whose best or second-best target was accepted by an EARLIER query of the same wave (one
// MATCH.ANY over the 2 x 16 target indices).  Everything before that query commits in parallel -- their targets are
// pairwise distinct by construction -- and the next wave starts at the conflicting query, which is then evaluated
// against the updated state.  The result is the reference's, query by query; a wave always commits at least its first.
// All threads first STAGE the problem in shared memory -- the candidate rows compacted back to back (block scan of the
// row lengths), the list of queries that have candidates, and for SearchForInitialization the per-target state
// vMatchedDistance / vnMatches21.  Problems whose rows exceed the shared-memory budget read the rows from global memory.
constexpr int kResolveWave = 16;
constexpr int kResolveThreads = 32 * kResolveWave;

// top-2 of one query's row under the current skip state; whole warp, result warp-uniform
__device__ __forceinline__ void resolve_row_top2(const unsigned* row, int c, int lane, bool init, const int* md, const uint8_t* taken,
                                                 unsigned& k1, unsigned& k2) {
    k1 = kNoKey; k2 = kNoKey;
    for (int j = lane; j < c; j += 32) {
        const unsigned e = row[j];
        const int idx = (int)(e >> kRowDistBits), dist = (int)(e & ((1u << kRowDistBits) - 1u));
        bool skip;
        if (init) skip = md[idx] <= dist;                              // :445-446
        else skip = taken[idx] != 0 || dist >= 256;                    // :88-90, :1404-1406, :210-211
        if (!skip) top2_push(k1, k2, ((unsigned)dist << 20) | (unsigned)j);
    }
    top2_warp_merge(k1, k2);
}

__global__ void __launch_bounds__(kResolveThreads) k_resolve(ResolveArgs a, int row_budget, int state_in_smem) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    __shared__ int s_hist[FBE_HISTO_LENGTH];
    __shared__ int s_ind[3];
    __shared__ int s_warp[kResolveThreads / 32];
    __shared__ int s_carry, s_nlist, s_base;
    __shared__ int s_bd[kResolveWave], s_bd2[kResolveWave], s_bi[kResolveWave], s_si[kResolveWave], s_l1[kResolveWave], s_l2[kResolveWave];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nq = a.nq[b], nt = a.nt[b];
    const size_t qb = (size_t)b * a.q_stride, tb = (size_t)b * a.t_stride;
    const bool init = a.mode == kResolveInit;

    int* s_off = reinterpret_cast<int*>(rs_smem);                       // [q_stride + 1] row starts in s_rows
    int* s_ql = s_off + a.q_stride + 1;                                 // [q_stride] queries that have candidates, in order
    int* s_md = s_ql + a.q_stride;                                      // [t_stride] (INIT, state_in_smem)
    int* s_m21 = s_md + (state_in_smem ? a.t_stride : 0);
    unsigned* s_rows = reinterpret_cast<unsigned*>(s_m21 + (state_in_smem ? a.t_stride : 0));
    int* md = state_in_smem ? s_md : a.matched_dist + tb;
    int* m21 = state_in_smem ? s_m21 : a.match21 + tb;
    const uint8_t* taken = init ? nullptr : a.taken + tb;

    // ---- stage: row offsets (exclusive block scan of cnt), rows, state -----------------------------------------------
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nq; base += kResolveThreads) {
        const int i = base + tid;
        const int v = i < nq ? a.cnt[qb + i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kResolveThreads / 32; ++w) { const int c = s_warp[w]; if (w < wid) before += c; tot += c; }
        const int carry = s_carry;
        if (i < nq) s_off[i] = carry + before + inc - v;
        __syncthreads();
        if (tid == 0) s_carry = carry + tot;
        __syncthreads();
    }
    const int total_rows = s_carry;
    if (tid == 0) s_off[nq] = total_rows;
    const bool staged = total_rows <= row_budget;
    __syncthreads();                           // s_off[nq] visible
    if (wid == 0) {
        // list the queries that have candidates (most keypoints are not octave-0 queries)
        int n = 0;
        for (int qbase = 0; qbase < nq; qbase += 32) {
            const int q = qbase + lane;
            const bool has = q < nq && s_off[q + 1] > s_off[q];
            const unsigned bal = __ballot_sync(0xffffffffu, has);
            if (has) s_ql[n + __popc(bal & ((1u << lane) - 1u))] = q;
            n += __popc(bal);
        }
        if (lane == 0) { s_nlist = n; s_base = 0; }
    }
    __syncthreads();
    const int nlist = s_nlist;
    if (staged) {
        // copy the rows: 4 list entries per warp per pass, their first 64 entries loaded before anything is stored, so the
        // global loads of one pass are in flight together (row lengths and offsets come from shared memory)
        for (int e0 = wid * 4; e0 < nlist; e0 += kResolveWave * 4) {
            int o[4], c[4];
            unsigned v0[4], v1[4];
            const unsigned* row[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = e0 + u < nlist ? s_ql[e0 + u] : -1;
                o[u] = q >= 0 ? s_off[q] : 0;
                c[u] = q >= 0 ? s_off[q + 1] - o[u] : 0;
                row[u] = a.rows + (qb + max(q, 0)) * a.C;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v0[u] = lane < c[u] ? row[u][lane] : 0u;
                v1[u] = lane + 32 < c[u] ? row[u][lane + 32] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (lane < c[u]) s_rows[o[u] + lane] = v0[u];
                if (lane + 32 < c[u]) s_rows[o[u] + lane + 32] = v1[u];
                for (int j = lane + 64; j < c[u]; j += 32) s_rows[o[u] + j] = row[u][j];
            }
        }
    }
    if (init) {
        for (int i = tid; i < nt; i += kResolveThreads) { md[i] = INT_MAX; m21[i] = -1; }
        for (int i = tid; i < nq; i += kResolveThreads) a.matches12[qb + i] = -1;
    } else {
        for (int i = tid; i < nt; i += kResolveThreads) a.cur_mp[tb + i] = -1;
    }
    for (int i = tid; i < nq; i += kResolveThreads) a.q_hit[qb + i] = -1;
    if (tid < FBE_HISTO_LENGTH) s_hist[tid] = 0;
    __syncthreads();

    int nm = 0;                                // warp 0: per-lane count of accepts minus steals, summed at the end
    const int sentinel = init ? INT_MAX : 256;
    for (int base = 0; base < nlist;) {
        // ---- evaluate: warp w takes query base + w of the list against the current state ------------------------------
        const int li = base + wid;
        if (li < nlist) {
            const int qi = s_ql[li];
            const int o = s_off[qi], c = s_off[qi + 1] - o;
            const unsigned* row = staged ? (s_rows + o) : (a.rows + (qb + qi) * a.C);
            unsigned k1, k2;
            resolve_row_top2(row, c, lane, init, md, taken, k1, k2);
            if (lane == 0) {
                const int bi = k1 == kNoKey ? -1 : (int)(row[k1 & 0xFFFFFu] >> kRowDistBits);
                const int si = k2 == kNoKey ? -1 : (int)(row[k2 & 0xFFFFFu] >> kRowDistBits);
                s_bd[wid] = (int)(k1 >> 20);
                s_bd2[wid] = k2 == kNoKey ? sentinel : (int)(k2 >> 20);
                s_bi[wid] = bi; s_si[wid] = si;
                if (a.mode == kResolveMap) {
                    s_l1[wid] = bi < 0 ? -1 : a.t_kps[tb + bi].octave;
                    s_l2[wid] = si < 0 ? -1 : a.t_kps[tb + si].octave;
                }
            }
        }
        __syncthreads();
        // ---- commit: warp 0, lane t = query base + t; lanes 16..31 carry the second-best targets for the conflict match --
        if (wid == 0) {
            const int t = lane & (kResolveWave - 1);
            const int cnt = min(kResolveWave, nlist - base);
            const bool live = t < cnt;
            const int bi = live ? s_bi[t] : -1, si = live ? s_si[t] : -1;
            bool accept = false;
            int bestDist = 0;
            if (live && bi >= 0) {
                bestDist = s_bd[t];
                const int bestDist2 = s_bd2[t];
                switch (a.mode) {
                    case kResolveInit: accept = bestDist <= FBE_TH_LOW && (float)bestDist < __fmul_rn((float)bestDist2, a.nn_ratio); break;
                    case kResolveLast: accept = bestDist <= (a.th_dist > 0 ? a.th_dist : FBE_TH_HIGH); break;
                    case kResolveMap:
                        accept = bestDist <= FBE_TH_HIGH;
                        if (accept && s_l1[t] == s_l2[t] && (float)bestDist > __fmul_rn(a.nn_ratio, (float)bestDist2)) accept = false;
                        break;
                    default: accept = bestDist <= (a.th_dist > 0 ? a.th_dist : FBE_TH_LOW) && (float)bestDist < __fmul_rn(a.nn_ratio, (float)bestDist2); break;
                }
            }
            // An earlier accept of this query's BEST target always matters.  An earlier accept of its SECOND best only makes
            // bestDist2 larger (or changes the second's level): that can flip a ratio-test reject into an accept, never the
            // reverse, and in the map search it can flip the same-level ratio gate either way.
            bool sens = false;
            if (live && bi >= 0) {
                if (a.mode == kResolveMap) sens = bestDist <= FBE_TH_HIGH;
                else if (a.mode != kResolveLast) sens = !accept && bestDist <= (init ? FBE_TH_LOW : (a.th_dist > 0 ? a.th_dist : FBE_TH_LOW));
            }
            const int mine = lane < kResolveWave ? bi : si;
            const unsigned same = __match_any_sync(0xffffffffu, mine >= 0 ? mine : -1 - lane);     // absent targets match nobody
            const unsigned same2 = __shfl_down_sync(0xffffffffu, same, kResolveWave);              // ... of this query's second best
            const unsigned acc = __ballot_sync(0xffffffffu, lane < kResolveWave && accept);
            const bool conflict = lane < kResolveWave && live && (((same | (sens ? same2 : 0u)) & acc & ((1u << lane) - 1u)) != 0u);
            const unsigned cf = __ballot_sync(0xffffffffu, conflict);
            const int ncommit = cf ? __ffs(cf) - 1 : cnt;            // >= 1: lane 0 has nobody before it
            if (lane < ncommit && accept) {
                const int qi = s_ql[base + lane];
                if (init) {
                    const int old = m21[bi];
                    if (old >= 0) { a.matches12[qb + old] = -1; --nm; }          // steal (:464-468); `old` is from an earlier wave
                    a.matches12[qb + qi] = bi;
                    m21[bi] = qi;
                    md[bi] = bestDist;
                } else {
                    a.cur_mp[tb + bi] = a.q_src ? a.q_src[qb + qi] : qi;
                    if (!a.q_has_obs || a.q_has_obs[qb + qi]) a.taken[tb + bi] = 1;
                }
                ++nm;
                a.q_hit[qb + qi] = bi;         // every accept is one rotHist push (:474-484), stolen or not
            }
            if (lane == 0) s_base = base + ncommit;
        }
        __syncthreads();
        base = s_base;
    }
    // ---- epilogue on the whole CTA: rotation histogram of all pushes, the three maxima, removals, vnMatches12 -> prev_matched --
    __shared__ int s_nm;
    if (wid == 0) {
        nm = __reduce_add_sync(0xffffffffu, nm);
        if (lane == 0) s_nm = nm;
    }
    __syncthreads();
    if (a.check_ori && a.mode != kResolveMap) {
        for (int qi = tid; qi < nq; qi += kResolveThreads) {
            const int hit = a.q_hit[qb + qi];
            int bin = -1;
            if (hit >= 0) {
                const int src = a.q_src ? a.q_src[qb + qi] : qi;
                bin = rot_bin(a.q_kps[qb + src].angle, a.t_kps[tb + hit].angle);
                atomicAdd(&s_hist[bin], 1);
            }
            a.q_bin[qb + qi] = bin;
        }
        __syncthreads();
        if (tid == 0) { int i1, i2, i3; three_maxima(s_hist, i1, i2, i3); s_ind[0] = i1; s_ind[1] = i2; s_ind[2] = i3; }
        __syncthreads();
        const int i1 = s_ind[0], i2 = s_ind[1], i3 = s_ind[2];
        int dec = 0;
        for (int qi = tid; qi < nq; qi += kResolveThreads) {       // same thread -> same qi as above: it reads its own q_bin
            const int bin = a.q_bin[qb + qi];
            if (bin < 0 || bin == i1 || bin == i2 || bin == i3) continue;
            if (init) {
                if (a.matches12[qb + qi] >= 0) { a.matches12[qb + qi] = -1; ++dec; }
            } else {
                a.cur_mp[tb + a.q_hit[qb + qi]] = -2;     // assigned by this call, then removed: the reference writes NULL
                ++dec;
            }
        }
        dec = __reduce_add_sync(0xffffffffu, dec);
        if (lane == 0 && dec) atomicSub(&s_nm, dec);
        __syncthreads();
    }
    if (init && a.prev_matched) {
        for (int qi = tid; qi < nq; qi += kResolveThreads) {
            const int m = a.matches12[qb + qi];
            if (m >= 0) a.prev_matched[qb + qi] = make_float2(a.t_kps[tb + m].x, a.t_kps[tb + m].y);
        }
    }
    if (tid == 0) a.nmatches[b] = s_nm;
}

// ---- BirdviewMatch epilogue (src/ORBmatcher.cc:1700-1759): ratio test, orientation histogram, DMatch list -----------
__global__ void __launch_bounds__(256) k_bird_finish(BirdFinishArgs a) {
    __shared__ int s_hist[FBE_HISTO_LENGTH];
    __shared__ int s_ind[3];
    __shared__ int s_warp[8];
    __shared__ int s_nm, s_carry;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nq = a.nq[b];
    const size_t qb = (size_t)b * a.q_stride, tb = (size_t)b * a.t_stride;
    if (tid < FBE_HISTO_LENGTH) s_hist[tid] = 0;
    if (tid == 0) { s_nm = 0; s_carry = 0; }
    __syncthreads();
    int nm = 0;
    for (int i = tid; i < nq; i += 256) {
        const int bi = a.best_idx[qb + i], d1 = a.best_dist[qb + i], d2 = a.second_dist[qb + i];
        int m = -1, bin = -1;
        if (bi >= 0 && d1 <= FBE_TH_LOW) {
            if ((float)d1 < __fmul_rn((float)d2, a.nn_ratio)) { m = bi; ++nm; }
            if (a.check_ori) { bin = rot_bin(a.q_kps[qb + i].angle, a.t_kps[tb + bi].angle); atomicAdd(&s_hist[bin], 1); }   // Q7
        }
        a.matches12[qb + i] = m;
        a.q_bin[qb + i] = bin;
    }
    __syncthreads();
    if (a.check_ori) {
        if (tid == 0) { int i1, i2, i3; three_maxima(s_hist, i1, i2, i3); s_ind[0] = i1; s_ind[1] = i2; s_ind[2] = i3; }
        __syncthreads();
        const int i1 = s_ind[0], i2 = s_ind[1], i3 = s_ind[2];
        for (int i = tid; i < nq; i += 256) {
            const int bin = a.q_bin[qb + i];
            if (bin < 0 || bin == i1 || bin == i2 || bin == i3) continue;
            if (a.matches12[qb + i] >= 0) { a.matches12[qb + i] = -1; --nm; }
        }
    }
    if (nm) atomicAdd(&s_nm, nm);
    __syncthreads();
    // DMatch(i, matches12[i], best_dist[i]) for matches12[i] > 0, in index order (Q8)
    if (a.dmatches) {
        const int lane = tid & 31, wid = tid >> 5;
        for (int base = 0; base < nq; base += 256) {
            const int i = base + tid;
            const int m = i < nq ? a.matches12[qb + i] : -1;
            const bool f = m > 0;
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_warp[wid] = __popc(bal);
            __syncthreads();
            int before = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const int c = s_warp[w]; if (w < wid) before += c; tot += c; }
            const int carry = s_carry;
            if (f) {
                int* o = a.dmatches + (qb + carry + before + __popc(bal & ((1u << lane) - 1u))) * 3;
                o[0] = i; o[1] = m; o[2] = a.best_dist[qb + i];
            }
            __syncthreads();
            if (tid == 0) s_carry = carry + tot;
            __syncthreads();
        }
        if (tid == 0) a.n_dmatches[b] = s_carry;
    }
    if (tid == 0) a.nmatches[b] = s_nm;
}

// ---- BirdMapPointMatch first-pass epilogue (src/ORBmatcher.cc:1852-1862) --------------------------------------------
__global__ void k_map_finish(const int* __restrict__ best_idx, const int* __restrict__ best_dist, const int* __restrict__ second_dist,
                             int n, float nn_ratio, int th_dist, int* __restrict__ matches12, int* __restrict__ nmatches) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int ok = 0;
    if (i < n) {
        const int bi = best_idx[i], d1 = best_dist[i], d2 = second_dist[i];
        ok = bi >= 0 && d1 <= th_dist && (float)d1 < __fmul_rn((float)d2, nn_ratio);
        matches12[i] = ok ? bi : -1;
    }
    const int c = __reduce_add_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(nmatches, c);
}

// ---- SearchByBoW rows: query = key-frame feature, candidates = the frame's features under the same vocabulary node ----
__global__ void __launch_bounds__(256) k_bow_rows(const uint8_t* __restrict__ kf_desc, const uint8_t* __restrict__ f_desc,
                                                  const int* __restrict__ q_src, const int* __restrict__ q_beg,
                                                  const int* __restrict__ q_end, const int* __restrict__ f_items, int nq, int C,
                                                  unsigned* __restrict__ rows, int* __restrict__ cnt) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    const int lane = threadIdx.x & 31;
    uint32_t qd[8];
    load_desc(qd, kf_desc + (size_t)q_src[qi] * 32);
    const int beg = q_beg[qi], end = q_end[qi];
    for (int p = beg + lane; p < end; p += 32) {
        const int idx = f_items[p];
        rows[(size_t)qi * C + (p - beg)] = ((unsigned)idx << kRowDistBits) | (unsigned)hamming256(qd, f_desc + (size_t)idx * 32);
    }
    if (lane == 0) cnt[qi] = end - beg;
}

// ---- brute force top-2 (stress config C5): POPC-pipe bound --------------------------------------------------------------
// 128 queries per CTA, one per thread with its descriptor in registers; targets stream through shared memory in tiles
// (every lane reads the same target word -> broadcast, no bank conflicts).  blockIdx.y splits the target set so that
// the grid covers the 148 SMs; a merge kernel folds the partial top-2 keys.
constexpr int kBfThreads = 128, kBfTile = 256;
__global__ void __launch_bounds__(kBfThreads) k_bruteforce(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt,
                                                           int chunk, unsigned* __restrict__ partial) {
    __shared__ __align__(16) uint4 s_t[kBfTile * 2];
    const int qi = blockIdx.x * kBfThreads + threadIdx.x;
    uint32_t qd[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (qi < nq) load_desc(qd, q + (size_t)qi * 32);
    const int t0 = blockIdx.y * chunk, t1 = min(nt, t0 + chunk);
    unsigned k1 = kNoKey, k2 = kNoKey;
    for (int base = t0; base < t1; base += kBfTile) {
        const int m = min(kBfTile, t1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < m * 2; i += kBfThreads) s_t[i] = reinterpret_cast<const uint4*>(t + (size_t)base * 32)[i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
            const uint4 b0 = s_t[2 * j], b1 = s_t[2 * j + 1];
            const int d = __popc(qd[0] ^ b0.x) + __popc(qd[1] ^ b0.y) + __popc(qd[2] ^ b0.z) + __popc(qd[3] ^ b0.w) +
                          __popc(qd[4] ^ b1.x) + __popc(qd[5] ^ b1.y) + __popc(qd[6] ^ b1.z) + __popc(qd[7] ^ b1.w);
            // keep the two smallest keys with min/max only (no branches): k2' = min(k2, max(k1, k)), k1' = min(k1, k)
            const unsigned k = ((unsigned)d << 20) | (unsigned)(base + j);
            k2 = min(k2, max(k1, k));
            k1 = min(k1, k);
        }
    }
    if (qi < nq) {
        partial[((size_t)blockIdx.y * nq + qi) * 2] = k1;
        partial[((size_t)blockIdx.y * nq + qi) * 2 + 1] = k2;
    }
}
__global__ void k_bruteforce_merge(const unsigned* __restrict__ partial, int nq, int nchunks, int* __restrict__ best_idx,
                                   int* __restrict__ best_dist, int* __restrict__ second_dist) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    unsigned k1 = kNoKey, k2 = kNoKey;
    for (int c = 0; c < nchunks; ++c) {
        top2_push(k1, k2, partial[((size_t)c * nq + qi) * 2]);
        top2_push(k1, k2, partial[((size_t)c * nq + qi) * 2 + 1]);
    }
    best_idx[qi] = k1 == kNoKey ? -1 : (int)(k1 & 0xFFFFFu);
    best_dist[qi] = k1 == kNoKey ? 257 : (int)(k1 >> 20);
    second_dist[qi] = k2 == kNoKey ? 257 : (int)(k2 >> 20);
}

// ---- launch wrappers -------------------------------------------------------------------------------------------------
int launch_queries_from_kps(const fbe_keypoint* kps, const float2* pos, const uint8_t*, const int* n, int stride, int nb,
                            float window, float4* q, int2* lv, cudaStream_t st) {
    dim3 grid((stride + 255) / 256, nb);
    k_queries_from_kps<<<grid, 256, 0, st>>>(kps, pos, n, stride, window, q, lv);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int launch_window_rows(const FrameDev& f, const QueryDev& qs, int nb, int max_nq, bool incl, int C, unsigned* rows, int* cnt,
                       int* overflow, cudaStream_t st) {
    if (max_nq <= 0) return FBE_OK;
    dim3 grid((max_nq + 7) / 8, nb);
    k_window_rows<<<grid, 256, 0, st>>>(f, qs, incl, C, rows, cnt, overflow);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int launch_window_top2(const FrameDev& f, const QueryDev& qs, int nb, int max_nq, bool incl, int* best_idx, int* best_dist,
                       int* second_dist, cudaStream_t st) {
    if (max_nq <= 0) return FBE_OK;
    dim3 grid((max_nq + 7) / 8, nb);
    k_window_top2<false><<<grid, 256, 0, st>>>(f, qs, incl, ReprojGate(), best_idx, best_dist, second_dist);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int launch_window_top2_reproj(const FrameDev& f, const QueryDev& qs, int max_nq, const ReprojGate& rg, int* best_idx, int* best_dist,
                              int* second_dist, cudaStream_t st) {
    if (max_nq <= 0) return FBE_OK;
    dim3 grid((max_nq + 7) / 8, 1);
    k_window_top2<true><<<grid, 256, 0, st>>>(f, qs, true, rg, best_idx, best_dist, second_dist);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int launch_resolve(const ResolveArgs& a, int nb, cudaStream_t st) {
    // shared memory: row offsets + query list + (INIT) per-target state + as many staged row entries as fit in the rest
    const size_t kMax = 200 * 1024;
    size_t fixed = (size_t)(2 * a.q_stride + 1) * 4;
    int state_in_smem = 0;
    if (a.mode == kResolveInit && fixed + (size_t)a.t_stride * 8 <= kMax / 2) { state_in_smem = 1; fixed += (size_t)a.t_stride * 8; }
    if (fixed > kMax) { set_error("resolve: too many queries for the shared-memory offsets"); return FBE_E_UNSUPPORTED; }
    // typical rows are far below the capacity C; budget = 16 entries per query on average, capped by what is left
    size_t budget = std::min((kMax - fixed) / 4, (size_t)a.q_stride * 16 + 1024);
    const size_t smem = fixed + budget * 4;
    if (smem > 48 * 1024) FBE_CUDA(cudaFuncSetAttribute(k_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMax));
    k_resolve<<<nb, kResolveThreads, smem, st>>>(a, (int)budget, state_in_smem);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int launch_bird_finish(const BirdFinishArgs& a, int nb, cudaStream_t st) {
    k_bird_finish<<<nb, 256, 0, st>>>(a);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int launch_map_finish(const int* best_idx, const int* best_dist, const int* second_dist, int n, float nn_ratio, int th_dist,
                      int* matches12, int* nmatches, cudaStream_t st) {
    if (n <= 0) return FBE_OK;
    k_map_finish<<<(n + 255) / 256, 256, 0, st>>>(best_idx, best_dist, second_dist, n, nn_ratio, th_dist, matches12, nmatches);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int launch_bow_rows(const uint8_t* kf_desc, const uint8_t* f_desc, const int* q_src, const int* q_beg, const int* q_end,
                    const int* f_items, int nq, int C, unsigned* rows, int* cnt, cudaStream_t st) {
    if (nq <= 0) return FBE_OK;
    k_bow_rows<<<(nq + 7) / 8, 256, 0, st>>>(kf_desc, f_desc, q_src, q_beg, q_end, f_items, nq, C, rows, cnt);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int launch_bruteforce(const uint8_t* q, int nq, const uint8_t* t, int nt, unsigned* partial, int nchunks, int* best_idx,
                      int* best_dist, int* second_dist, cudaStream_t st) {
    if (nq <= 0) return FBE_OK;
    const int chunk = (((nt + nchunks - 1) / nchunks) + kBfTile - 1) / kBfTile * kBfTile;
    dim3 grid((nq + kBfThreads - 1) / kBfThreads, nchunks);
    k_bruteforce<<<grid, kBfThreads, 0, st>>>(q, nq, t, nt, std::max(chunk, kBfTile), partial);
    k_bruteforce_merge<<<(nq + 255) / 256, 256, 0, st>>>(partial, nq, nchunks, best_idx, best_dist, second_dist);
    count_launch(2);
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

// ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-307) -------------------------------------------------
// One warp per map point, one lane per row of its N x N distance matrix.  The row median sorted_row[(N-1)/2] is found
// without storing the row: bisection on the value v in [0, 256] with count(dist <= v) recomputed per step (9 passes of
// N Hamming distances; N is the number of observations of a map point, tens at most, and the descriptors of one point are
// 32 x N contiguous bytes that stay in L1).  The first row with the least median wins (`median < BestMedian`, :293).
__global__ void __launch_bounds__(128) k_distinctive(const uint8_t* __restrict__ desc, const int* __restrict__ start, int npts,
                                                     int* __restrict__ best, int* __restrict__ best_median) {
    const int p = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (p >= npts) return;
    const int s = start[p], N = start[p + 1] - s;
    if (N <= 0) {
        if (lane == 0) { best[p] = -1; if (best_median) best_median[p] = 0; }
        return;
    }
    const int need = (N - 1) / 2 + 1;                              // rank (int)(0.5*(N-1)) -> that many values <= median
    const uint8_t* d0 = desc + (size_t)s * 32;
    unsigned key = 0xFFFFFFFFu;
    for (int i = lane; i < N; i += 32) {
        uint32_t a[8];
        load_desc(a, d0 + (size_t)i * 32);
        int lo = 0, hi = 256;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int cnt = 0;
            for (int j = 0; j < N; ++j) cnt += hamming256(a, d0 + (size_t)j * 32) <= mid;
            if (cnt >= need) hi = mid; else lo = mid + 1;
        }
        key = min(key, ((unsigned)lo << 20) | (unsigned)i);
    }
    key = __reduce_min_sync(0xffffffffu, key);
    if (lane == 0) { best[p] = (int)(key & 0xFFFFFu); if (best_median) best_median[p] = (int)(key >> 20); }
}

int launch_distinctive(const uint8_t* desc, const int* start, int npts, int* best, int* best_median, cudaStream_t st) {
    if (npts <= 0) return FBE_OK;
    k_distinctive<<<(npts + 3) / 4, 128, 0, st>>>(desc, start, npts, best, best_median);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

// ---- ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:658-824) -------------------------------------------------
// The reference never sets vbMatched2, so queries do not interact: each key-frame-1 feature takes, among the key-frame-2
// features of its vocabulary node that clear the filters (no map point, dist <= TH_LOW, not within 100*scale of the
// epipole for monocular pairs, CheckDistEpipolarLine :141-158), the one with the smallest distance, the LAST one on ties
// (`dist>bestDist` skips, equality replaces).  One warp per query, lanes over the node list, one packed-key warp minimum.
__global__ void __launch_bounds__(256) k_tri_rows(const TriArgs a) {
    const int q = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (q >= a.nq) return;
    const int i1 = a.q_src[q], beg = a.q_beg[q], end = a.q_end[q];
    const fbe_keypoint kp1 = a.kps1[i1];
    const bool st1 = a.stereo1[i1] != 0;
    uint32_t d1[8];
    load_desc(d1, a.desc1 + (size_t)i1 * 32);
    // epipolar line in the second image l = x1' F12 (:144-146)
    const float la = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, a.F[0]), __fmul_rn(kp1.y, a.F[3])), a.F[6]);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, a.F[1]), __fmul_rn(kp1.y, a.F[4])), a.F[7]);
    const float lc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, a.F[2]), __fmul_rn(kp1.y, a.F[5])), a.F[8]);
    const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
    unsigned key = 0xFFFFFFFFu;
    for (int p = beg + lane; p < end; p += 32) {
        const int i2 = a.items2[p];
        if (a.skip2[i2]) continue;
        const int dist = hamming256(d1, a.desc2 + (size_t)i2 * 32);
        if (dist > FBE_TH_LOW) continue;
        const fbe_keypoint kp2 = a.kps2[i2];
        if (!st1 && !a.stereo2[i2]) {
            const float dx = __fsub_rn(a.ex, kp2.x), dy = __fsub_rn(a.ey, kp2.y);
            if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, a.scale[kp2.octave])) continue;
        }
        const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, kp2.x), __fmul_rn(lb, kp2.y)), lc);
        if (den == 0.0f) continue;
        const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
        if (!((double)dsqr < 3.84 * (double)a.sigma2[kp2.octave])) continue;
        key = min(key, ((unsigned)dist << 20) | (0xFFFFFu - (unsigned)(p - beg)));
    }
    key = __reduce_min_sync(0xffffffffu, key);
    if (lane == 0) {
        int best = -1, bin = -1;
        if (key != 0xFFFFFFFFu) {
            best = a.items2[beg + (int)(0xFFFFFu - (key & 0xFFFFFu))];
            if (a.check_ori) bin = rot_bin(kp1.angle, a.kps2[best].angle);
        }
        a.q_best[q] = best; a.q_bin[q] = bin;
    }
}

// rotation-consistency prune (:779-797) and the scatter into vMatches12
__global__ void __launch_bounds__(256) k_tri_finish(const TriArgs a) {
    __shared__ int s_hist[FBE_HISTO_LENGTH], s_ind[3], s_nm;
    const int tid = threadIdx.x;
    if (tid < FBE_HISTO_LENGTH) s_hist[tid] = 0;
    if (tid == 0) { s_nm = 0; s_ind[0] = s_ind[1] = s_ind[2] = -1; }
    __syncthreads();
    if (a.check_ori)
        for (int q = tid; q < a.nq; q += 256)
            if (a.q_best[q] >= 0) atomicAdd(&s_hist[a.q_bin[q]], 1);
    __syncthreads();
    if (a.check_ori && tid == 0) { int i1, i2, i3; three_maxima(s_hist, i1, i2, i3); s_ind[0] = i1; s_ind[1] = i2; s_ind[2] = i3; }
    __syncthreads();
    int nm = 0;
    for (int q = tid; q < a.nq; q += 256) {
        const int best = a.q_best[q];
        if (best < 0) continue;
        const int bin = a.q_bin[q];
        if (a.check_ori && bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) continue;
        a.matches12[a.q_src[q]] = best;
        ++nm;
    }
    if (nm) atomicAdd(&s_nm, nm);
    __syncthreads();
    if (tid == 0) *a.nmatches = s_nm;
}

int launch_triangulation(const TriArgs& a, cudaStream_t st) {
    if (a.nq <= 0) return FBE_OK;
    k_tri_rows<<<(a.nq + 7) / 8, 256, 0, st>>>(a);
    k_tri_finish<<<1, 256, 0, st>>>(a);
    count_launch(2);
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
