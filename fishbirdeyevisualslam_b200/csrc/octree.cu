// DistributeOctTree (src/ORBextractor.cc:539-763, DivideNode :481-537) as a round-parallel replay.
//
// One CTA per (image, level).  The reference walks a std::list and splits nodes one at a time; the selection it
// produces depends on list order, on the order children are created and on a sort by (size, node address).  Two facts
// make an exact parallel replay possible:
//   1. after the roots the only list insertion is push_front, so the list is always "live nodes by creation sequence,
//      newest first".  We keep the live nodes in an ARRAY in ascending creation order (array index == rank); the list
//      front is the last array element.
//   2. within one sweep (or one refinement round) the nodes to split are fixed up front and their splits are mutually
//      independent; only the ORDER of processing matters (it numbers the children, and in refinement it decides where
//      the "stop at N" cut falls).  Sweep order = list order = descending array index.  Refinement order = (size,
//      creation sequence) descending -- the canonical reading of the reference's sort over (size, pointer) pairs with a
//      monotonic allocator (the oracle's verbatim reference build runs on a bump arena for the same reason).
// Keys never move: each key carries the array index of its node, child sizes come from shared-memory atomics, and
// "first key in candidate order with the maximum response" is an atomicMax over (score, -candidate index).
#include <algorithm>
#include "fbe_internal.cuh"

namespace fbe {

constexpr int kOctThreads = 256;    // measured: 512 -> 2.913, 384 -> 2.866, 256 -> 2.833, 192 -> 2.842, 128 -> 2.882 ms per step (the replay is barrier-bound)

// bounds relative to (16,16): UL=(x0,y0) BR=(x1,y1).  A node is final (bNoMore) iff cnt == 1.  pd = depth << 26 | path
// index: the node's position in the data-independent quad-tree geometry (root r, then one base-4 digit per split).
struct __align__(16) OctNode { short x0, y0, x1, y1; int cnt; int pd; };
constexpr int kPdShift = 26;
__device__ __forceinline__ bool nomore(const OctNode& nd) { return nd.cnt == 1; }

__device__ __forceinline__ int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct BlockScan {
    int* warp_sums;   // [kOctThreads/32 + 1] shared
    // exclusive scan of one int per thread; returns exclusive prefix, writes block total to *total
    __device__ __forceinline__ int exclusive(int v, int& total) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __syncthreads();                       // protect warp_sums from the previous use
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kOctThreads / 32; ++w) {
            int c = warp_sums[w];
            if (w < wid) before += c;
            tot += c;
        }
        total = tot;
        return before + inc - v;
    }
};

// quadrant of a key inside a node: n1=0 (left,top) n2=1 (right,top) n3=2 (left,bottom) n4=3 (right,bottom).
// The reference compares float key coordinates with int bounds (:515-525); coordinates are integral here.
__device__ __forceinline__ int quadrant(uint32_t key, const OctNode& nd) {
    const int x = key_x(key) - 16, y = key_y(key) - 16;
    const int mx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2)
    const int my = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
    return (x < mx ? 0 : 1) + (y < my ? 0 : 2);
}

// Walk of all keys with the (key, node) loads of 4 iterations issued before any of them is used: the arrays live in
// global memory (a level can hold > 16k candidates) and every pass of the replay is bound by that latency.
template <class F>
__device__ __forceinline__ void for_keys(const uint32_t* __restrict__ keys, const uint32_t* knode, int nk, F&& f) {
    int k = threadIdx.x;
    for (; k + 3 * kOctThreads < nk; k += 4 * kOctThreads) {
        const uint32_t a0 = keys[k], a1 = keys[k + kOctThreads], a2 = keys[k + 2 * kOctThreads], a3 = keys[k + 3 * kOctThreads];
        const uint32_t n0 = knode[k], n1 = knode[k + kOctThreads], n2 = knode[k + 2 * kOctThreads], n3 = knode[k + 3 * kOctThreads];
        f(k, a0, n0); f(k + kOctThreads, a1, n1); f(k + 2 * kOctThreads, a2, n2); f(k + 3 * kOctThreads, a3, n3);
    }
    for (; k < nk; k += kOctThreads) f(k, keys[k], knode[k]);
}

// Walk of one array with the loads of 8 iterations in flight (same reason as for_keys)
template <class F>
__device__ __forceinline__ void for_each8(const uint32_t* arr, int nk, F&& f) {
    int k = threadIdx.x;
    for (; k + 7 * kOctThreads < nk; k += 8 * kOctThreads) {
        uint32_t a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = arr[k + u * kOctThreads];
#pragma unroll
        for (int u = 0; u < 8; ++u) f(k + u * kOctThreads, a[u]);
    }
    for (; k < nk; k += kOctThreads) f(k, arr[k]);
}

// counter[idx] += 1 for every calling lane, with the lanes of a warp that hit the same counter folded into ONE shared-memory
// atomic (MATCH.ANY): the first rounds have a handful of nodes, so 16k keys would otherwise serialise on <= 8 addresses.
__device__ __forceinline__ void agg_inc(int* counter, int idx) {
    const unsigned mask = __activemask();
    const unsigned peers = __match_any_sync(mask, idx);
    if ((threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&counter[idx], __popc(peers));
}

__device__ __forceinline__ int nonempty4(const int* c4) { return (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0); }

// ---- histogram mode ----------------------------------------------------------------------------------------------
// The geometry of the quad tree (node bounds at every depth) does not depend on the data: only which children exist
// and in which order nodes are processed does.  One pass therefore records, for every key, the depth-D cell of the
// geometry it falls into (histogram of nini * 4^D counters); the counters of all coarser depths are sums of those.
// While no node deeper than D - 1 has to be split, every round of the replay takes its child sizes from the histogram
// and the keys are not touched at all; they are assigned to their final nodes in one pass at the end (or at the
// moment a node at depth D must be split, after which the replay continues with per-key passes).
// depth of the histogram for a level with nini roots and N wanted nodes: one more than the depth at which the full tree
// has N nodes (so that the refinement round is covered too), reduced until the counters fit 16k ints
__host__ __device__ inline int oct_hist_depth(int nini, int N) {
    int d = 0;
    while ((nini << (2 * d)) < N && d < 8) ++d;
    int D = d + 1;
    while (D > 0 && nini * (((1 << (2 * (D + 1))) - 1) / 3) > 16384) --D;
    return D;
}
__host__ __device__ inline int oct_hist_entries(int nini, int D) { return D > 0 ? nini * (((1 << (2 * (D + 1))) - 1) / 3) : 0; }

__device__ __forceinline__ int hist_offset(int nini, int d) { return nini * (((1 << (2 * d)) - 1) / 3); }   // sum_{j<d} nini*4^j

// walks key (x,y) down the geometry from root r; calls visit(depth, path) at every depth 0..D until it returns true
template <class F>
__device__ __forceinline__ void descend(int x, int y, int r, float hx, int H, int D, F&& visit) {
    int x0 = (int)__fmul_rn(hx, (float)r), x1 = (int)__fmul_rn(hx, (float)(r + 1)), y0 = 0, y1 = H;
    int path = r;
    for (int d = 0;; ++d) {
        if (visit(d, path) || d == D) return;
        const int mx = x0 + ((x1 - x0 + 1) >> 1), my = y0 + ((y1 - y0 + 1) >> 1);
        int q = 0;
        if (x < mx) x1 = mx; else { x0 = mx; q = 1; }
        if (y < my) y1 = my; else { y0 = my; q += 2; }
        path = 4 * path + q;
    }
}

__device__ __forceinline__ int root_of(uint32_t key, float hx, int nini) {
    const int x = key_x(key) - 16;
    const int r = (int)__fdiv_rn((float)x, hx);         // vpIniNodes[kp.pt.x/hX]
    return min(max(r, 0), nini - 1);
}

// assigns every key to the live node it belongs to: table[off(depth) + path] = node index for live nodes, -1 elsewhere
__device__ void materialise_knode(uint32_t* knode, int nk, const OctNode* cur, int n, int nini, int D, int* table, int table_n) {
    const int tid = threadIdx.x;
    for (int i = tid; i < table_n; i += kOctThreads) table[i] = -1;
    __syncthreads();
    for (int i = tid; i < n; i += kOctThreads) {
        const int pd = cur[i].pd;
        table[hist_offset(nini, pd >> kPdShift) + (pd & ((1 << kPdShift) - 1))] = i;
    }
    __syncthreads();
    // the histogram pass parked every key's depth-D path in knode[k]; its path at depth d is that value >> 2 (D - d)
    // (path = 4 * path + quadrant per level), so the walk down the geometry is not repeated
    for_each8(knode, nk, [&](int k, uint32_t leaf) {
        int node = 0;
        for (int d = 0; d <= D; ++d) {
            const int t = table[hist_offset(nini, d) + ((int)leaf >> (2 * (D - d)))];
            if (t >= 0) { node = t; break; }
        }
        knode[k] = (uint32_t)node;
    });
    __syncthreads();
}

// One axis of descend(): the depth-D path digits contributed by coordinate v inside [lo, hi) (bit `bit` of every base-4 digit)
__device__ __forceinline__ int descend_axis(int v, int lo, int hi, int D, int bit) {
    int path = 0;
    for (int d = 0; d < D; ++d) {
        const int m = lo + ((hi - lo + 1) >> 1);
        int q = 0;
        if (v < m) hi = m; else { lo = m; q = bit; }
        path = 4 * path + q;
    }
    return path;
}

// The replay proper.  keys/knode: nk entries.  Returns the number of live nodes; `cur` points at the final array.
// All pointers may be shared or global memory.
__device__ int octree_replay(const uint32_t* __restrict__ keys, uint32_t* knode, int nk, int nini, float hx, int W, int H,
                             int N, int cap, OctNode* nodesA, OctNode* nodesB, int* cnt4, int* cnt4b, int* newidx, int* splitf,
                             int* order, unsigned long long* sortbuf, BlockScan& bs, int* s_ctl, OctNode** out_nodes, int* hist, int D) {
    const int tid = threadIdx.x;
    OctNode* cur = nodesA;
    OctNode* nxt = nodesB;
    bool hist_mode = D > 0;                       // D == 0: no histogram (too large for the scratch), per-key passes only
    const int hist_n = hist_mode ? hist_offset(nini, D + 1) : 0;

    // ---- roots (:543-585) -------------------------------------------------------------------------------------
    if (hist_mode) {
        for (int i = tid; i < hist_n; i += kOctThreads) hist[i] = 0;
        __syncthreads();
        int* hD = hist + hist_offset(nini, D);
        // The x and y decisions of descend() are independent, so the depth-D path of (x, y) is xtab[x] + ytab[y]: root and x digits of
        // every column, y digits of every row, tabulated once per CTA (W + H one-axis walks instead of one two-axis walk per key).
        // The tables borrow cnt4b, which the replay does not touch before its first round; levels whose tables do not fit walk per key.
        uint16_t* xtab = reinterpret_cast<uint16_t*>(cnt4b);
        uint16_t* ytab = xtab + W;
        const bool tabs = W > 0 && 2 * (W + H) <= 16 * cap && (nini << (2 * D)) <= 65536;
        if (tabs) {
            for (int i = tid; i < W + H; i += kOctThreads) {
                if (i < W) {
                    const int r = min(max((int)__fdiv_rn((float)i, hx), 0), nini - 1);            // root_of()
                    xtab[i] = (uint16_t)((r << (2 * D)) + descend_axis(i, (int)__fmul_rn(hx, (float)r), (int)__fmul_rn(hx, (float)(r + 1)), D, 1));
                } else {
                    ytab[i - W] = (uint16_t)descend_axis(i - W, 0, H, D, 2);
                }
            }
            __syncthreads();
        }
        if (tabs) {
            for_each8(keys, nk, [&](int k, uint32_t key) {
                const int leaf = (int)xtab[key_x(key) - 16] + (int)ytab[key_y(key) - 16];
                knode[k] = (uint32_t)leaf;         // parked for materialise_knode()
                agg_inc(hD, leaf);
            });
        } else {
            for (int k = tid; k < nk; k += kOctThreads) {
                const uint32_t key = keys[k];
                int leaf = 0;
                descend(key_x(key) - 16, key_y(key) - 16, root_of(key, hx, nini), hx, H, D, [&](int d, int path) { leaf = path; return false; });
                knode[k] = (uint32_t)leaf;
                agg_inc(hD, leaf);
            }
        }
        __syncthreads();
        for (int d = D - 1; d >= 0; --d) {        // coarser depths = sums of their four children
            int* hd = hist + hist_offset(nini, d);
            const int* hc = hist + hist_offset(nini, d + 1);
            for (int i = tid; i < (nini << (2 * d)); i += kOctThreads) hd[i] = hc[4 * i] + hc[4 * i + 1] + hc[4 * i + 2] + hc[4 * i + 3];
            __syncthreads();
        }
        for (int r = tid; r < nini; r += kOctThreads) cnt4[r] = hist[r];
    } else {
        for (int r = tid; r < nini; r += kOctThreads) cnt4[r] = 0;
        __syncthreads();
        for (int k = tid; k < nk; k += kOctThreads) {
            const int r = root_of(keys[k], hx, nini);
            knode[k] = (uint32_t)r;
            agg_inc(cnt4, r);
        }
    }
    __syncthreads();
    // array position of root r = number of non-empty roots with a larger index (root 0 is the list front = last)
    for (int r = tid; r < nini; r += kOctThreads) {
        int pos = 0;
        for (int r2 = r + 1; r2 < nini; ++r2) pos += cnt4[r2] > 0;
        newidx[r] = pos;
        if (cnt4[r] > 0) {
            OctNode nd;
            nd.x0 = (short)(int)__fmul_rn(hx, (float)r);
            nd.x1 = (short)(int)__fmul_rn(hx, (float)(r + 1));
            nd.y0 = 0; nd.y1 = (short)H;
            nd.cnt = cnt4[r];
            nd.pd = r;                                   // depth 0, path = root index
            cur[pos] = nd;
        }
    }
    if (tid == 0) {
        int n = 0;
        for (int r = 0; r < nini; ++r) n += cnt4[r] > 0;
        s_ctl[0] = n;
    }
    __syncthreads();
    int n = s_ctl[0];
    if (!hist_mode) {
        for (int k = tid; k < nk; k += kOctThreads) knode[k] = (uint32_t)newidx[knode[k]];
        __syncthreads();
        // ---- child sizes of every splittable root; later rounds get theirs from the re-homing pass of the round before ----
        for (int i = tid; i < 4 * n; i += kOctThreads) cnt4[i] = 0;
        __syncthreads();
        for_keys(keys, knode, nk, [&](int, uint32_t key, uint32_t node) {
            const OctNode nd = cur[node];
            if (!nomore(nd)) agg_inc(cnt4, 4 * (int)node + quadrant(key, nd));
        });
    }
    __syncthreads();

    bool refine = false;
    while (true) {
        const int prev = n;
        if (hist_mode) {
            // child sizes from the histogram -- unless a node that must be split already sits at depth D
            if (tid == 0) s_ctl[3] = 0;
            __syncthreads();
            bool deep = false;
            for (int i = tid; i < n; i += kOctThreads) {
                const OctNode nd = cur[i];
                deep |= !nomore(nd) && (nd.pd >> kPdShift) >= D;
            }
            if (deep) s_ctl[3] = 1;
            __syncthreads();
            if (s_ctl[3]) {
                hist_mode = false;               // leave histogram mode: keys get their nodes, then per-key counting
                materialise_knode(knode, nk, cur, n, nini, D, hist, hist_n);
                for (int i = tid; i < 4 * n; i += kOctThreads) cnt4[i] = 0;
                __syncthreads();
                for_keys(keys, knode, nk, [&](int, uint32_t key, uint32_t node) {
                    const OctNode nd = cur[node];
                    if (!nomore(nd)) agg_inc(cnt4, 4 * (int)node + quadrant(key, nd));
                });
            } else {
                for (int i = tid; i < n; i += kOctThreads) {
                    const OctNode nd = cur[i];
                    if (!nomore(nd)) {
                        const int* hc = hist + hist_offset(nini, (nd.pd >> kPdShift) + 1) + 4 * (nd.pd & ((1 << kPdShift) - 1));
                        cnt4[4 * i] = hc[0]; cnt4[4 * i + 1] = hc[1]; cnt4[4 * i + 2] = hc[2]; cnt4[4 * i + 3] = hc[3];
                    }
                }
            }
            __syncthreads();
        }

        // ---- processing order of the splittable nodes --------------------------------------------------------
        int m = 0;   // number of splittable nodes
        if (!refine) {
            // sweep: list order = descending array index
            for (int base = 0; base < n; base += kOctThreads) {
                const int j = base + tid;            // j-th node from the back
                const int i = n - 1 - j;
                const int f = (j < n) && !nomore(cur[i]);
                int tot;
                const int ex = bs.exclusive(f, tot);
                if (f) order[m + ex] = i;
                m += tot;
            }
            __syncthreads();
            for (int i = tid; i < n; i += kOctThreads) splitf[i] = !nomore(cur[i]);
        } else {
            // refinement: (size, creation rank) descending; stop once the live count reaches N (:685-732)
            for (int base = 0; base < n; base += kOctThreads) {
                const int i = base + tid;
                const int f = (i < n) && !nomore(cur[i]);
                int tot;
                const int ex = bs.exclusive(f, tot);
                if (f) sortbuf[m + ex] = ((unsigned long long)(unsigned)cur[i].cnt << 32) | (unsigned)i;
                m += tot;
            }
            // descending sort of the m unique (size, rank) keys.  Small m (the normal case: a few hundred splittable nodes):
            // rank sort -- every thread counts the keys greater than its own, ONE barrier; large m: bitonic network.
            unsigned long long* sorted = sortbuf;
            if (m <= 2 * kOctThreads) {
                __syncthreads();
                unsigned long long* out = reinterpret_cast<unsigned long long*>(cnt4b);      // free until the re-homing pass
                for (int t = tid; t < m; t += kOctThreads) {
                    const unsigned long long mine = sortbuf[t];
                    int rank = 0;
                    for (int u = 0; u < m; ++u) rank += sortbuf[u] > mine;
                    out[rank] = mine;
                }
                sorted = out;
                __syncthreads();
            } else {
                const int mp = pow2_ceil(max(m, 1));
                for (int j = m + tid; j < mp; j += kOctThreads) sortbuf[j] = 0ull;
                __syncthreads();
                for (int k2 = 2; k2 <= mp; k2 <<= 1) {
                    for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                        for (int t = tid; t < mp; t += kOctThreads) {
                            const int p = t ^ j2;
                            if (p > t) {
                                const unsigned long long a = sortbuf[t], c = sortbuf[p];
                                const bool desc = (t & k2) == 0;      // descending overall
                                if (desc ? (a < c) : (a > c)) { sortbuf[t] = c; sortbuf[p] = a; }
                            }
                        }
                        __syncthreads();
                    }
                }
            }
            for (int i = tid; i < n; i += kOctThreads) splitf[i] = 0;
            __syncthreads();
            // running live count after each split in order; everything up to and including the first split that
            // reaches N is performed
            int carry = 0;
            if (tid == 0) s_ctl[1] = m;          // cut position (exclusive end of performed splits)
            __syncthreads();
            for (int base = 0; base < m; base += kOctThreads) {
                const int j = base + tid;
                int gain = 0, i = 0;
                if (j < m) { i = (int)(sorted[j] & 0xffffffffu); gain = nonempty4(cnt4 + 4 * i) - 1; }
                int tot;
                const int ex = bs.exclusive(gain, tot);
                const int live_before = n + carry + ex;          // live count before this split
                if (j < m) {
                    order[j] = i;
                    if (live_before < N) {
                        splitf[i] = 1;
                        if (live_before + gain >= N) atomicMin(&s_ctl[1], j + 1);
                    }
                }
                carry += tot;
            }
            __syncthreads();
            m = s_ctl[1];     // only the first m entries of `order` are split (prefix property of the cut)
            __syncthreads();
        }

        // ---- new array = unsplit nodes (ascending) ++ children in creation order ------------------------------
        int nsurv = 0;
        for (int base = 0; base < n; base += kOctThreads) {
            const int i = base + tid;
            const int f = (i < n) && !splitf[i];
            int tot;
            const int ex = bs.exclusive(f, tot);
            if (f) { newidx[i] = nsurv + ex; nxt[nsurv + ex] = cur[i]; }
            nsurv += tot;
        }
        int nchild = 0, nexp_local = 0;
        for (int base = 0; base < m; base += kOctThreads) {
            const int j = base + tid;
            int i = 0, c = 0;
            if (j < m) { i = order[j]; c = nonempty4(cnt4 + 4 * i); }
            int tot;
            const int ex = bs.exclusive(c, tot);
            if (j < m) {
                int pos = nsurv + nchild + ex;
                newidx[i] = pos;                         // child base of a split node
                if (pos + c <= cap) {
                    const OctNode nd = cur[i];
                    const int mx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), my = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cq = cnt4[4 * i + q];
                        if (cq > 0) {
                            OctNode ch;
                            ch.x0 = (q & 1) ? (short)mx : nd.x0;  ch.x1 = (q & 1) ? nd.x1 : (short)mx;
                            ch.y0 = (q & 2) ? (short)my : nd.y0;  ch.y1 = (q & 2) ? nd.y1 : (short)my;
                            ch.cnt = cq;
                            ch.pd = (((nd.pd >> kPdShift) + 1) << kPdShift) | (4 * (nd.pd & ((1 << kPdShift) - 1)) + q);
                            nxt[pos++] = ch;
                            nexp_local += cq > 1;
                        }
                    }
                }
            }
            nchild += tot;
        }
        const int n2 = nsurv + nchild;
        if (tid == 0) s_ctl[2] = 0;
        if (n2 <= cap && !hist_mode)
            for (int i = tid; i < 4 * n2; i += kOctThreads) cnt4b[i] = 0;
        __syncthreads();
        if (nexp_local) atomicAdd(&s_ctl[2], nexp_local);
        // ---- re-home the keys and, in the same pass, count the child sizes of their NEW nodes (next round's input) ----
        if (n2 <= cap && !hist_mode) {
            for_keys(keys, knode, nk, [&](int k, uint32_t key, uint32_t node) {
                const int i = (int)node;
                int dst = newidx[i];
                if (splitf[i]) {
                    const int q = quadrant(key, cur[i]);
                    const int* c4 = cnt4 + 4 * i;
                    dst += (q > 0 && c4[0] > 0) + (q > 1 && c4[1] > 0) + (q > 2 && c4[2] > 0);
                    knode[k] = (uint32_t)dst;
                } else if (dst != i) {
                    knode[k] = (uint32_t)dst;
                }
                const OctNode nd = nxt[dst];
                if (!nomore(nd)) agg_inc(cnt4b, 4 * dst + quadrant(key, nd));
            });
        }
        __syncthreads();
        const int nexp = s_ctl[2];
        __syncthreads();
        OctNode* t = cur; cur = nxt; nxt = t;
        if (!hist_mode) { int* tc = cnt4; cnt4 = cnt4b; cnt4b = tc; }
        if (n2 > cap) { n = -1; break; }          // cannot happen (see DESIGN.md bound); guarded anyway
        n = n2;
        if (n >= N || n == prev) break;
        if (!refine && n + 3 * nexp > N) refine = true;
    }
    if (hist_mode && n > 0) materialise_knode(knode, nk, cur, n, nini, D, hist, hist_n);
    *out_nodes = cur;
    return n;
}

__global__ void __launch_bounds__(kOctThreads, 4) k_octree(const Plan* __restrict__ plan, Workspace ws, int use_smem, int level_stride) {
    extern __shared__ __align__(16) uint8_t dyn[];
    __shared__ int s_warp[kOctThreads / 32 + 1];
    __shared__ int s_ctl[4];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    // grid = (images, levels): CTAs are dispatched x-fastest, so every level-0 CTA (the longest: a level's work shrinks by
    // 1 / scaleFactor^2 per level) starts first and the short high levels fill the tail (longest-processing-time-first)
    const int l = blockIdx.y, b = blockIdx.x;
    const LevelGeom g = plan->lv[l];
    BlockScan bs{s_warp};

    uint32_t* keys = ws.keys + (size_t)b * plan->slots_total + g.slot_base;
    uint32_t* knode = ws.key_node + (size_t)b * plan->slots_total + g.slot_base;
    const uint32_t* slots = ws.slots + (size_t)b * plan->slots_total + g.slot_base;
    const int* ccount = ws.cell_count + (size_t)b * plan->ncells_total + g.cell_base;
    const int ncells = g.ncols * g.nrows;

    // ---- vToDistributeKeys: concatenate the cells in row-major order ----------------------------------------------
    // (cell offsets by a block scan, parked in the not-yet-used knode array; then one warp per cell copies coalesced)
    int nk = 0;
    for (int base = 0; base < ncells; base += kOctThreads) {
        const int c = base + tid;
        const int cnt = c < ncells ? ccount[c] : 0;
        int tot;
        const int ex = bs.exclusive(cnt, tot);
        if (c < ncells) knode[c] = (uint32_t)(nk + ex);
        nk += tot;
    }
    __syncthreads();
    {
        // four cells per warp per pass; count, offset and the first 32 slots of each are loaded before any of them is used (a slot
        // beyond the count holds stale keys, never out-of-bounds memory), so a pass costs one global-memory latency, not two per cell
        const int lane = tid & 31, wid = tid >> 5;
        const bool lane_in_cap = lane < g.cell_cap;
        for (int c0 = wid * 4; c0 < ncells; c0 += (kOctThreads / 32) * 4) {
            int cnt[4], off[4];
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = min(c0 + u, ncells - 1);
                cnt[u] = c0 + u < ncells ? ccount[c] : 0;
                off[u] = (int)knode[c];
                v[u] = lane_in_cap ? slots[(size_t)c * g.cell_cap + lane] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (lane < cnt[u]) keys[off[u] + lane] = v[u];
                if (cnt[u] > 32) {
                    const uint32_t* src = slots + (size_t)(c0 + u) * g.cell_cap;
                    for (int i = lane + 32; i < cnt[u]; i += 32) keys[off[u] + i] = src[i];
                }
            }
        }
    }
    __syncthreads();

    // ---- scratch carve-up -----------------------------------------------------------------------------------------
    const int cap = g.node_cap;
    // global fallback: every level owns `level_stride` bytes of the per-image scratch
    uint8_t* base_ptr = use_smem ? dyn : ws.oct_scratch + (size_t)b * ws.oct_scratch_bytes + (size_t)l * level_stride;
    OctNode* nodesA = reinterpret_cast<OctNode*>(base_ptr);
    OctNode* nodesB = nodesA + cap;
    unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(nodesB + cap);
    int* cnt4 = reinterpret_cast<int*>(sortbuf + pow2_ceil(cap));
    int* cnt4b = cnt4 + 4 * cap;
    int* newidx = cnt4b + 4 * cap;
    int* splitf = newidx + cap;
    int* order = splitf + cap;
    int* hist = order + cap;
    const int D = oct_hist_depth(g.nini, g.nfeat);

    const int W = g.w - 2 * kEdge + 6, H = g.h - 2 * kEdge + 6;   // maxBorder - minBorder
    OctNode* fin = nullptr;
    int n = octree_replay(keys, knode, nk, g.nini, g.hx, W, H, g.nfeat, cap, nodesA, nodesB, cnt4, cnt4b, newidx, splitf, order,
                          sortbuf, bs, s_ctl, &fin, hist, D);
    int* level_n = ws.level_n + (size_t)b * FBE_MAX_LEVELS + l;
    if (n < 0) {
        if (tid == 0) { *level_n = 0; atomicOr(ws.status + b, 1); }
        return;
    }
    // ---- best key of every live node (:742-760): max response, first in candidate order -----------------------------
    unsigned* best = reinterpret_cast<unsigned*>(cnt4);
    for (int i = tid; i < n; i += kOctThreads) best[i] = 0u;
    __syncthreads();
    for_keys(keys, knode, nk, [&](int k, uint32_t key, uint32_t node) {
        atomicMax(&best[node], ((unsigned)key_s(key) << 24) | (0xFFFFFFu - (unsigned)k));
    });
    __syncthreads();
    uint32_t* sel = ws.sel + (size_t)b * plan->kp_cap_total + g.kp_base;
    for (int j = tid; j < n && j < g.kp_cap; j += kOctThreads) {
        const unsigned v = best[n - 1 - j];                  // list front = last array element
        sel[j] = keys[0xFFFFFFu - (v & 0xFFFFFFu)];
    }
    if (tid == 0) *level_n = min(n, g.kp_cap);
}

// Stand-alone entry used by fbe_debug_octree (adversarial parity tests): same replay, scratch in global memory,
// output = selected candidate INDICES in reference output order.
__global__ void __launch_bounds__(kOctThreads) k_octree_debug(const uint32_t* keys, uint32_t* knode, int nk, int nini, float hx,
                                                              int H, int N, int cap, uint8_t* scratch, uint32_t* sel_idx, int* n_out) {
    __shared__ int s_warp[kOctThreads / 32 + 1];
    __shared__ int s_ctl[4];
    const int tid = threadIdx.x;
    BlockScan bs{s_warp};
    OctNode* nodesA = reinterpret_cast<OctNode*>(scratch);
    OctNode* nodesB = nodesA + cap;
    unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(nodesB + cap);
    int* cnt4 = reinterpret_cast<int*>(sortbuf + pow2_ceil(cap));
    int* cnt4b = cnt4 + 4 * cap;
    int* newidx = cnt4b + 4 * cap;
    int* splitf = newidx + cap;
    int* order = splitf + cap;
    int* hist = order + cap;
    OctNode* fin = nullptr;
    int n = octree_replay(keys, knode, nk, nini, hx, /*W: unknown here, per-key walk*/ 0, H, N, cap, nodesA, nodesB, cnt4, cnt4b, newidx, splitf, order, sortbuf, bs, s_ctl, &fin,
                          hist, oct_hist_depth(nini, N));
    if (n < 0) { if (tid == 0) *n_out = -1; return; }
    unsigned* best = reinterpret_cast<unsigned*>(cnt4);
    for (int i = tid; i < n; i += kOctThreads) best[i] = 0u;
    __syncthreads();
    for_keys(keys, knode, nk, [&](int k, uint32_t key, uint32_t node) {
        atomicMax(&best[node], ((unsigned)key_s(key) << 24) | (0xFFFFFFu - (unsigned)k));
    });
    __syncthreads();
    for (int j = tid; j < n; j += kOctThreads) sel_idx[j] = 0xFFFFFFu - (best[n - 1 - j] & 0xFFFFFFu);
    if (tid == 0) *n_out = n;
}

static size_t oct_level_bytes(int cap, int nini, int N) {
    int p2 = 1;
    while (p2 < cap) p2 <<= 1;
    const size_t hist = (size_t)oct_hist_entries(nini, oct_hist_depth(nini, N)) * 4;
    return (((size_t)cap * (16 + 16 + 16 + 16 + 4 + 4 + 4) + (size_t)p2 * 8 + hist + 64) + 127) & ~(size_t)127;
}

static size_t oct_level_stride(const Plan& hp) {
    size_t need = 0;
    for (int l = 0; l < hp.nlevels; ++l) need = std::max(need, oct_level_bytes(hp.lv[l].node_cap, hp.lv[l].nini, hp.lv[l].nfeat));
    return need;
}

size_t octree_scratch_bytes(const Plan& hp) { return oct_level_stride(hp) * hp.nlevels + 1024; }

size_t octree_debug_scratch_bytes(int cap, int nini, int N) { return oct_level_bytes(cap, nini, N); }

int launch_octree_debug(const uint32_t* d_keys, uint32_t* d_knode, int nk, int nini, float hx, int H, int nfeat, int cap,
                        uint8_t* d_scratch, uint32_t* d_sel_idx, int* d_n, cudaStream_t st) {
    k_octree_debug<<<1, kOctThreads, 0, st>>>(d_keys, d_knode, nk, nini, hx, H, nfeat, cap, d_scratch, d_sel_idx, d_n);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

int launch_octree(const Plan& hp, const Plan* dp, const Workspace& ws, int nimg, cudaStream_t st) {
    const size_t need = oct_level_stride(hp);
    const int use_smem = need <= 160 * 1024;
    size_t smem = use_smem ? need : 0;
    if (smem > 48 * 1024) FBE_CUDA(cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nimg, hp.nlevels);
    FBE_CUDA(launch_dep(k_octree, grid, dim3(kOctThreads), smem, st, dp, ws, use_smem, (int)need));
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe
