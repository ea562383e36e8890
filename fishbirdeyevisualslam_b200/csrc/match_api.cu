// C-ABI of the matcher family (include/fbe_cabi.h): host buffers in/out, all searches on the device.
// The host side only marshals: it flattens frames to arrays, builds the per-query (centre, radius, level range)
// records with the reference's fp32 arithmetic, and copies results back.  No Hamming distance is computed on the host.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>
#include "match_kernels.cuh"

using namespace fbe;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return FBE_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        FBE_CUDA(cudaMalloc(&p, want));
        cap = want;
        return FBE_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// pinned host staging for results: a device->host copy into the caller's pageable arrays is staged by the driver and costs a
// round trip each; results land here with ONE synchronise per search and are copied out by the host
struct HostPin {
    uint8_t* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return FBE_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        FBE_CUDA(cudaMallocHost(&p, want));
        cap = want;
        return FBE_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct FrameBufs {
    DevBuf kps, desc, n, cell, start, items;
    void release() { kps.release(); desc.release(); n.release(); cell.release(); start.release(); items.release(); }
};

// Device-resident frames of the matcher (SURVEY §8b `*_dev` intent for the per-frame drop-in chain): Tracking searches the
// same frame several times in a row (SearchForInitialization retries, SearchByProjection with two radii, local-map search
// after the last-frame search, bird + front on the same Frame ...).  The last few frames handed in through fbe_frame_view stay
// on the device together with their CSR grid; a view is recognised by CONTENT (a host shadow copy is compared byte for byte:
// ~5 us for 2000 keypoints, cheaper than three pageable uploads + the grid build), never by pointer, so a recycled
// std::vector address can not alias a stale entry.
struct CachedFrame {
    FrameBufs b;
    std::vector<uint8_t> shadow_kps, shadow_desc;
    int n = -1;
    bool has_grid = false;
    float min_x = 0, min_y = 0, inv_w = 0, inv_h = 0;
    int gcols = 0, grows = 0;
    unsigned long long used = 0;
};
constexpr int kFrameCache = 4;

}  // namespace

struct fbe_matcher {
    float nn_ratio;
    int check_ori;
    int device;
    cudaStream_t stream = nullptr;
    int row_cap = 128;
    FrameBufs fa, fb;                          // target frame / auxiliary source frame (array-based entry points)
    CachedFrame cache[kFrameCache];            // frames given as fbe_frame_view
    unsigned long long tick = 0, cache_hits = 0, cache_misses = 0;
    DevBuf q, lv, qdesc, nq, rows, cnt, misc, i0, i1, i2, i3, i4, u0, u1, f0, partial;
    HostPin hres;                              // [0, 64): flags + counts, [64, ...): the search's result arrays
};

namespace {

#define FBE_TRY(expr) do { int _rc = (expr); if (_rc != FBE_OK) return _rc; } while (0)

int upload(DevBuf& b, const void* src, size_t bytes, cudaStream_t st) {
    FBE_TRY(b.ensure(std::max<size_t>(bytes, 16)));
    if (bytes) FBE_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
    return FBE_OK;
}

// keypoints + descriptors of a frame on the device, with its CSR grid when asked for: served from the matcher's frame cache
// when the same content was seen recently, uploaded (and grid-built) otherwise
int upload_frame(fbe_matcher* m, FrameBufs& /*unused: cached frames own their buffers*/, const fbe_frame_view* v, bool with_grid, FrameDev& out) {
    if (!v || v->n < 0 || (v->n > 0 && (!v->kps || !v->desc))) return FBE_E_INVALID;
    if (with_grid && (v->gcols <= 0 || v->grows <= 0)) return FBE_E_INVALID;
    const int n = v->n, stride = std::max(n, 1);
    const size_t kb = (size_t)n * sizeof(fbe_keypoint), db = (size_t)n * 32;
    const unsigned long long tick = ++m->tick;
    CachedFrame* e = nullptr;
    for (CachedFrame& c : m->cache)
        if (c.n == n && (n == 0 || (std::memcmp(c.shadow_kps.data(), v->kps, kb) == 0 && std::memcmp(c.shadow_desc.data(), v->desc, db) == 0))) { e = &c; break; }
    if (e) {
        ++m->cache_hits;
    } else {
        ++m->cache_misses;
        e = &m->cache[0];
        for (CachedFrame& c : m->cache) if (c.used < e->used) e = &c;       // least recently used (the other frame of this call is the newest)
        e->n = -1; e->has_grid = false;
        FBE_TRY(upload(e->b.kps, v->kps, kb, m->stream));
        FBE_TRY(upload(e->b.desc, v->desc, db, m->stream));
        FBE_TRY(upload(e->b.n, &v->n, sizeof(int), m->stream));
        // the caller's arrays cannot change before this (synchronous) entry point returns: the shadow equals what was uploaded
        e->shadow_kps.assign(reinterpret_cast<const uint8_t*>(v->kps), reinterpret_cast<const uint8_t*>(v->kps) + kb);
        e->shadow_desc.assign(v->desc, v->desc + db);
        e->n = n;
    }
    e->used = tick;
    out = FrameDev();
    out.kps = e->b.kps.as<fbe_keypoint>(); out.desc = e->b.desc.as<uint8_t>(); out.n = e->b.n.as<int>();
    out.kp_stride = stride;
    out.min_x = v->min_x; out.min_y = v->min_y; out.inv_w = v->inv_w; out.inv_h = v->inv_h;
    out.gcols = v->gcols; out.grows = v->grows;
    if (with_grid) {
        const bool same = e->has_grid && e->min_x == v->min_x && e->min_y == v->min_y && e->inv_w == v->inv_w && e->inv_h == v->inv_h &&
                          e->gcols == v->gcols && e->grows == v->grows;
        if (!same) {
            e->has_grid = false;
            FBE_TRY(e->b.cell.ensure((size_t)stride * 4));
            FBE_TRY(e->b.start.ensure((size_t)(v->gcols * v->grows + 1) * 4));
            FBE_TRY(e->b.items.ensure((size_t)stride * 4));
            FBE_TRY(launch_grid_build(out.kps, out.n, stride, 1, v->min_x, v->min_y, v->inv_w, v->inv_h, v->gcols, v->grows,
                                      e->b.cell.as<int>(), e->b.start.as<int>(), e->b.items.as<int>(), m->stream));
            e->min_x = v->min_x; e->min_y = v->min_y; e->inv_w = v->inv_w; e->inv_h = v->inv_h; e->gcols = v->gcols; e->grows = v->grows;
            e->has_grid = true;
        }
        out.start = e->b.start.as<int>(); out.items = e->b.items.as<int>();
    }
    return FBE_OK;
}

int upload_queries(fbe_matcher* m, const std::vector<float4>& q, const std::vector<int2>& lv, const uint8_t* desc, int nq,
                   QueryDev& out) {
    const int stride = std::max(nq, 1);
    FBE_TRY(upload(m->q, q.data(), (size_t)nq * sizeof(float4), m->stream));
    FBE_TRY(upload(m->lv, lv.data(), (size_t)nq * sizeof(int2), m->stream));
    FBE_TRY(upload(m->qdesc, desc, (size_t)nq * 32, m->stream));
    FBE_TRY(upload(m->nq, &nq, sizeof(int), m->stream));
    out.q = m->q.as<float4>(); out.lv = m->lv.as<int2>(); out.desc = m->qdesc.as<uint8_t>(); out.nq = m->nq.as<int>();
    out.stride = stride;
    return FBE_OK;
}

// rows + sequential resolve with automatic growth of the per-query row capacity.  before_resolve(attempt) restores the state the
// resolve kernel modifies (a retry must start from the caller's data again); enqueue_results() queues the device->host copies of
// the search's outputs into m->hres behind the resolve kernel, so that one synchronise serves flags, counts and results.
template <class Setup, class Results>
int rows_and_resolve(fbe_matcher* m, const FrameDev& tf, const QueryDev& qs, int nq, bool incl, ResolveArgs a, Setup&& before_resolve,
                     Results&& enqueue_results) {
    FBE_TRY(m->hres.ensure(64));
    for (int attempt = 0; attempt < 8; ++attempt) {
        const int C = m->row_cap;
        FBE_TRY(m->rows.ensure((size_t)std::max(nq, 1) * C * 4));
        FBE_TRY(m->cnt.ensure((size_t)std::max(nq, 1) * 4));
        FBE_TRY(m->misc.ensure(64));
        FBE_CUDA(cudaMemsetAsync(m->misc.p, 0, 64, m->stream));
        FBE_TRY(launch_window_rows(tf, qs, 1, nq, incl, C, m->rows.as<unsigned>(), m->cnt.as<int>(), m->misc.as<int>(), m->stream));
        FBE_TRY(before_resolve(attempt));
        a.C = C; a.rows = m->rows.as<unsigned>(); a.cnt = m->cnt.as<int>();
        a.nq = qs.nq; a.q_stride = qs.stride; a.t_stride = tf.kp_stride; a.nt = tf.n; a.t_kps = tf.kps;
        a.nmatches = m->misc.as<int>() + 1;
        FBE_TRY(launch_resolve(a, 1, m->stream));
        FBE_TRY(enqueue_results());
        FBE_CUDA(cudaMemcpyAsync(m->hres.p, m->misc.p, 8, cudaMemcpyDeviceToHost, m->stream));
        FBE_CUDA(cudaStreamSynchronize(m->stream));
        const int* h = reinterpret_cast<const int*>(m->hres.p);
        if (!h[0]) return h[1];           // >= 0 : nmatches
        m->row_cap *= 2;                  // a window held more candidates than the row capacity: redo, larger
    }
    set_error("candidate rows overflow");
    return FBE_E_CAPACITY;
}

}  // namespace

extern "C" {

int fbe_matcher_create(float nn_ratio, int32_t check_orientation, int32_t device, fbe_matcher** out) {
    if (!out) return FBE_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device: this library has no CPU path"); return FBE_E_CUDA; }
    FBE_CUDA(cudaSetDevice(device));
    fbe_matcher* m = new (std::nothrow) fbe_matcher();
    if (!m) return FBE_E_INVALID;
    m->nn_ratio = nn_ratio; m->check_ori = check_orientation != 0; m->device = device;
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { delete m; set_error("stream"); return FBE_E_CUDA; }
    *out = m;
    return FBE_OK;
}

int fbe_matcher_destroy(fbe_matcher* m) {
    if (!m) return FBE_E_INVALID;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    m->fa.release(); m->fb.release();
    for (CachedFrame& c : m->cache) c.b.release();
    m->hres.release();
    for (DevBuf* b : {&m->q, &m->lv, &m->qdesc, &m->nq, &m->rows, &m->cnt, &m->misc, &m->i0, &m->i1, &m->i2, &m->i3, &m->i4, &m->u0, &m->u1, &m->f0, &m->partial}) b->release();
    cudaStreamDestroy(m->stream);
    delete m;
    return FBE_OK;
}

int fbe_matcher_cache_stats(const fbe_matcher* m, uint64_t* hits, uint64_t* misses) {
    if (!m) return FBE_E_INVALID;
    if (hits) *hits = m->cache_hits;
    if (misses) *misses = m->cache_misses;
    return FBE_OK;
}

int fbe_search_for_initialization(fbe_matcher* m, const fbe_frame_view* f1, const fbe_frame_view* f2, float* prev_matched,
                                  int32_t* matches12, int32_t window_size, int32_t* nmatches) {
    if (!m || !f1 || !f2 || !nmatches || (f1->n > 0 && (!prev_matched || !matches12))) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(m->device));
    *nmatches = 0;
    const int n1 = f1->n, n2 = f2->n;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return FBE_OK;
    FrameDev F1, F2;
    FBE_TRY(upload_frame(m, m->fb, f1, false, F1));
    FBE_TRY(upload_frame(m, m->fa, f2, true, F2));
    FBE_TRY(upload(m->f0, prev_matched, (size_t)n1 * 8, m->stream));
    FBE_TRY(m->hres.ensure(64 + (size_t)n1 * 12));
    FBE_TRY(m->q.ensure((size_t)n1 * sizeof(float4)));
    FBE_TRY(m->lv.ensure((size_t)n1 * sizeof(int2)));
    FBE_TRY(launch_queries_from_kps(F1.kps, m->f0.as<float2>(), nullptr, F1.n, n1, 1, (float)window_size, m->q.as<float4>(),
                                    m->lv.as<int2>(), m->stream));
    QueryDev qs{m->q.as<float4>(), m->lv.as<int2>(), F1.desc, F1.n, n1};
    FBE_TRY(m->i0.ensure((size_t)n2 * 4)); FBE_TRY(m->i1.ensure((size_t)n2 * 4));
    FBE_TRY(m->i2.ensure((size_t)n1 * 4)); FBE_TRY(m->i3.ensure((size_t)n1 * 4)); FBE_TRY(m->i4.ensure((size_t)n1 * 4));
    ResolveArgs a{};
    a.mode = kResolveInit; a.q_kps = F1.kps; a.q_src = nullptr; a.q_has_obs = nullptr;
    a.nn_ratio = m->nn_ratio; a.check_ori = m->check_ori;
    a.matched_dist = m->i0.as<int>(); a.match21 = m->i1.as<int>(); a.matches12 = m->i2.as<int>();
    a.prev_matched = m->f0.as<float2>(); a.q_bin = m->i3.as<int>(); a.q_hit = m->i4.as<int>();
    // a retry after a row overflow must start from the caller's vbPrevMatched again (the first attempt still has the upload above)
    int rc = rows_and_resolve(m, F2, qs, n1, true, a,
        [&](int attempt) { return attempt ? upload(m->f0, prev_matched, (size_t)n1 * 8, m->stream) : FBE_OK; },
        [&]() {
            FBE_CUDA(cudaMemcpyAsync(m->hres.p + 64, m->i2.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, m->stream));
            FBE_CUDA(cudaMemcpyAsync(m->hres.p + 64 + (size_t)n1 * 4, m->f0.p, (size_t)n1 * 8, cudaMemcpyDeviceToHost, m->stream));
            return FBE_OK;
        });
    if (rc < 0) return rc;
    *nmatches = rc;
    std::memcpy(matches12, m->hres.p + 64, (size_t)n1 * 4);
    std::memcpy(prev_matched, m->hres.p + 64 + (size_t)n1 * 4, (size_t)n1 * 8);
    return FBE_OK;
}

int fbe_birdview_match(fbe_matcher* m, const fbe_keypoint* ref_kps, const uint8_t* ref_desc, int32_t n_ref,
                       const fbe_frame_view* cur, int32_t window_size, int32_t* dmatches, int32_t* n_dmatches, int32_t* nmatches) {
    if (!m || !cur || !nmatches || !n_dmatches || n_ref < 0 || (n_ref > 0 && (!ref_kps || !ref_desc || !dmatches))) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(m->device));
    *nmatches = 0; *n_dmatches = 0;
    if (n_ref == 0 || cur->n == 0) return FBE_OK;
    fbe_frame_view rv = *cur;
    rv.kps = ref_kps; rv.desc = ref_desc; rv.n = n_ref;
    FrameDev R, Cf;
    FBE_TRY(upload_frame(m, m->fb, &rv, false, R));
    FBE_TRY(upload_frame(m, m->fa, cur, true, Cf));
    FBE_TRY(m->q.ensure((size_t)n_ref * sizeof(float4)));
    FBE_TRY(m->lv.ensure((size_t)n_ref * sizeof(int2)));
    FBE_TRY(launch_queries_from_kps(R.kps, nullptr, nullptr, R.n, n_ref, 1, (float)window_size, m->q.as<float4>(), m->lv.as<int2>(), m->stream));
    QueryDev qs{m->q.as<float4>(), m->lv.as<int2>(), R.desc, R.n, n_ref};
    FBE_TRY(m->i0.ensure((size_t)n_ref * 4)); FBE_TRY(m->i1.ensure((size_t)n_ref * 4)); FBE_TRY(m->i2.ensure((size_t)n_ref * 4));
    FBE_TRY(m->i3.ensure((size_t)n_ref * 4)); FBE_TRY(m->i4.ensure((size_t)n_ref * 12)); FBE_TRY(m->misc.ensure(64));
    FBE_TRY(launch_window_top2(Cf, qs, 1, n_ref, false, m->i0.as<int>(), m->i1.as<int>(), m->i2.as<int>(), m->stream));
    BirdFinishArgs a{};
    a.best_idx = m->i0.as<int>(); a.best_dist = m->i1.as<int>(); a.second_dist = m->i2.as<int>();
    a.nq = R.n; a.q_stride = n_ref; a.q_kps = R.kps; a.t_kps = Cf.kps; a.t_stride = Cf.kp_stride;
    a.nn_ratio = m->nn_ratio; a.check_ori = m->check_ori;
    a.matches12 = m->i3.as<int>(); a.dmatches = m->i4.as<int>(); a.n_dmatches = m->misc.as<int>(); a.nmatches = m->misc.as<int>() + 1;
    FBE_TRY(m->partial.ensure((size_t)n_ref * 4));
    a.q_bin = m->partial.as<int>();
    FBE_TRY(launch_bird_finish(a, 1, m->stream));
    int h[2];
    FBE_CUDA(cudaMemcpyAsync(h, m->misc.p, 8, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaStreamSynchronize(m->stream));
    *n_dmatches = h[0]; *nmatches = h[1];
    if (h[0] > 0) FBE_CUDA(cudaMemcpy(dmatches, m->i4.p, (size_t)h[0] * 12, cudaMemcpyDeviceToHost));
    return FBE_OK;
}

int fbe_bird_map_point_match(fbe_matcher* m, const float* mp_pix, const uint8_t* mp_desc, int32_t n_mp, const fbe_frame_view* cur,
                             int32_t window_size, int32_t* matches12, int32_t* nmatches) {
    if (!m || !cur || !nmatches || n_mp < 0 || (n_mp > 0 && (!mp_pix || !mp_desc || !matches12))) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(m->device));
    *nmatches = 0;
    for (int i = 0; i < n_mp; ++i) matches12[i] = -1;
    if (n_mp == 0 || cur->n == 0) return FBE_OK;
    std::vector<float4> q(n_mp);
    std::vector<int2> lv(n_mp, make_int2(-1, -1));
    for (int i = 0; i < n_mp; ++i) {
        const bool skip = std::isnan(mp_pix[2 * i]);
        q[i] = make_float4(skip ? 0.f : mp_pix[2 * i], skip ? 0.f : mp_pix[2 * i + 1], skip ? -1.f : (float)window_size, 0.f);
    }
    FrameDev Cf;
    QueryDev qs;
    FBE_TRY(upload_frame(m, m->fa, cur, true, Cf));
    FBE_TRY(upload_queries(m, q, lv, mp_desc, n_mp, qs));
    FBE_TRY(m->i0.ensure((size_t)n_mp * 4)); FBE_TRY(m->i1.ensure((size_t)n_mp * 4)); FBE_TRY(m->i2.ensure((size_t)n_mp * 4));
    FBE_TRY(m->i3.ensure((size_t)n_mp * 4)); FBE_TRY(m->misc.ensure(64));
    FBE_CUDA(cudaMemsetAsync(m->misc.p, 0, 64, m->stream));
    FBE_TRY(launch_window_top2(Cf, qs, 1, n_mp, false, m->i0.as<int>(), m->i1.as<int>(), m->i2.as<int>(), m->stream));
    FBE_TRY(launch_map_finish(m->i0.as<int>(), m->i1.as<int>(), m->i2.as<int>(), n_mp, m->nn_ratio, FBE_TH_LOW, m->i3.as<int>(),
                              m->misc.as<int>(), m->stream));
    FBE_CUDA(cudaMemcpyAsync(matches12, m->i3.p, (size_t)n_mp * 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaMemcpyAsync(nmatches, m->misc.p, 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaStreamSynchronize(m->stream));
    return FBE_OK;
}

static int projection_search(fbe_matcher* m, int mode, const fbe_frame_view* cur, const std::vector<float4>& q,
                             const std::vector<int2>& lv, const uint8_t* qdesc, const fbe_keypoint* q_kps, int nq,
                             const uint8_t* cur_taken, const uint8_t* q_has_obs, const int* q_src, int check_ori,
                             int32_t* cur_mp, int32_t* nmatches, int th_dist = 0) {
    const int nt = cur->n;
    FrameDev Cf;
    QueryDev qs;
    FBE_TRY(upload_frame(m, m->fa, cur, true, Cf));
    FBE_TRY(upload_queries(m, q, lv, qdesc, nq, qs));
    std::vector<uint8_t> taken(nt, 0);
    if (cur_taken) std::memcpy(taken.data(), cur_taken, nt);
    std::vector<uint8_t> obs(std::max(nq, 1), 1);
    if (q_has_obs) std::memcpy(obs.data(), q_has_obs, nq);
    FBE_TRY(upload(m->u1, obs.data(), (size_t)nq, m->stream));
    FBE_TRY(m->i0.ensure((size_t)nt * 4)); FBE_TRY(m->i3.ensure((size_t)std::max(nq, 1) * 4)); FBE_TRY(m->i4.ensure((size_t)std::max(nq, 1) * 4));
    if (q_kps) FBE_TRY(upload(m->fb.kps, q_kps, (size_t)nq * sizeof(fbe_keypoint), m->stream));
    if (q_src) FBE_TRY(upload(m->i1, q_src, (size_t)nq * 4, m->stream));
    ResolveArgs a{};
    a.mode = mode; a.q_kps = q_kps ? m->fb.kps.as<fbe_keypoint>() : nullptr; a.q_src = q_src ? m->i1.as<int>() : nullptr;
    a.q_has_obs = m->u1.as<uint8_t>(); a.nn_ratio = m->nn_ratio; a.check_ori = check_ori; a.th_dist = th_dist;
    a.cur_mp = m->i0.as<int>(); a.q_bin = m->i3.as<int>(); a.q_hit = m->i4.as<int>();
    FBE_TRY(m->u0.ensure((size_t)nt));
    a.taken = m->u0.as<uint8_t>();
    FBE_TRY(m->hres.ensure(64 + (size_t)nt * 4));
    int rc = rows_and_resolve(m, Cf, qs, nq, true, a,
        [&](int) { return upload(m->u0, taken.data(), (size_t)nt, m->stream); },      // fresh `taken` state on every attempt
        [&]() {
            FBE_CUDA(cudaMemcpyAsync(m->hres.p + 64, m->i0.p, (size_t)nt * 4, cudaMemcpyDeviceToHost, m->stream));
            return FBE_OK;
        });
    if (rc < 0) return rc;
    *nmatches = rc;
    std::memcpy(cur_mp, m->hres.p + 64, (size_t)nt * 4);
    return FBE_OK;
}

int fbe_search_by_projection_last(fbe_matcher* m, const fbe_frame_view* cur, const fbe_keypoint* last_kps, const float* last_proj,
                                  const uint8_t* last_mp_desc, int32_t n_last, const float* scale_factors, int32_t nlevels,
                                  const uint8_t* cur_taken, const uint8_t* last_has_obs, float th, int32_t* cur_mp, int32_t* nmatches) {
    if (!m || !cur || !nmatches || n_last < 0 || !scale_factors || (cur->n > 0 && !cur_mp) ||
        (n_last > 0 && (!last_kps || !last_proj || !last_mp_desc))) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(m->device));
    *nmatches = 0;
    for (int k = 0; k < cur->n; ++k) cur_mp[k] = -1;
    if (n_last == 0 || cur->n == 0) return FBE_OK;
    std::vector<float4> q(n_last);
    std::vector<int2> lv(n_last);
    for (int i = 0; i < n_last; ++i) {
        const int o = last_kps[i].octave;
        if (o < 0 || o >= nlevels) return FBE_E_INVALID;
        const bool skip = std::isnan(last_proj[2 * i]);
        const float radius = th * scale_factors[o];          // :1382
        q[i] = make_float4(skip ? 0.f : last_proj[2 * i], skip ? 0.f : last_proj[2 * i + 1], skip ? -1.f : radius, 0.f);
        lv[i] = make_int2(o - 1, o + 1);                    // mono branch :1391
    }
    return projection_search(m, kResolveLast, cur, q, lv, last_mp_desc, last_kps, n_last, cur_taken, last_has_obs, nullptr,
                             m->check_ori, cur_mp, nmatches);
}

// Shared body of the relocalisation and loop-closing SearchByProjection overloads: single best unmatched candidate in
// th * scale[predicted level], levels [pl-1, pl+level_up], accepted below th_dist; every assignment blocks its keypoint.
static int search_by_projection_kf(fbe_matcher* m, const fbe_frame_view* cur, const fbe_keypoint* q_kps, const float* proj,
                                   const int32_t* level, const uint8_t* mp_desc, int32_t n_mp, const float* scale_factors,
                                   int32_t nlevels, const uint8_t* cur_taken, float th, int th_dist, int level_up, int check_ori,
                                   int32_t* cur_mp, int32_t* nmatches) {
    if (!m || !cur || !nmatches || n_mp < 0 || !scale_factors || th_dist <= 0 || (cur->n > 0 && !cur_mp) ||
        (n_mp > 0 && (!proj || !level || !mp_desc)) || (check_ori && n_mp > 0 && !q_kps)) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(m->device));
    *nmatches = 0;
    for (int k = 0; k < cur->n; ++k) cur_mp[k] = -1;
    if (n_mp == 0 || cur->n == 0) return FBE_OK;
    std::vector<float4> q(n_mp);
    std::vector<int2> lv(n_mp);
    for (int i = 0; i < n_mp; ++i) {
        const int pl = level[i];
        if (pl < 0 || pl >= nlevels) return FBE_E_INVALID;
        const bool skip = std::isnan(proj[2 * i]);
        const float radius = th * scale_factors[pl];             // :1527 / :355
        q[i] = make_float4(skip ? 0.f : proj[2 * i], skip ? 0.f : proj[2 * i + 1], skip ? -1.f : radius, 0.f);
        lv[i] = make_int2(pl - 1, pl + level_up);
    }
    return projection_search(m, kResolveLast, cur, q, lv, mp_desc, check_ori ? q_kps : nullptr, n_mp, cur_taken, nullptr, nullptr,
                             check_ori, cur_mp, nmatches, th_dist);
}

int fbe_search_by_projection_reloc(fbe_matcher* m, const fbe_frame_view* cur, const fbe_keypoint* kf_kps, const float* mp_proj,
                                   const int32_t* mp_level, const uint8_t* mp_desc, int32_t n_mp, const float* scale_factors,
                                   int32_t nlevels, const uint8_t* cur_taken, float th, int32_t orb_dist, int32_t* cur_mp,
                                   int32_t* nmatches) {
    if (!m) return FBE_E_INVALID;
    return search_by_projection_kf(m, cur, kf_kps, mp_proj, mp_level, mp_desc, n_mp, scale_factors, nlevels, cur_taken, th, orb_dist,
                                   1, m->check_ori, cur_mp, nmatches);
}

int fbe_search_by_projection_loop(fbe_matcher* m, const fbe_frame_view* kf, const float* mp_proj, const int32_t* mp_level,
                                  const uint8_t* mp_desc, int32_t n_mp, const float* scale_factors, int32_t nlevels,
                                  const uint8_t* kf_matched, int32_t th, int32_t* kf_mp, int32_t* nmatches) {
    return search_by_projection_kf(m, kf, nullptr, mp_proj, mp_level, mp_desc, n_mp, scale_factors, nlevels, kf_matched, (float)th,
                                   FBE_TH_LOW, 0, 0, kf_mp, nmatches);
}

int fbe_search_by_projection_map(fbe_matcher* m, const fbe_frame_view* cur, const float* scale_factors, int32_t nlevels,
                                 const float* mp_proj, const int32_t* mp_level, const float* mp_viewcos, const uint8_t* mp_desc,
                                 int32_t n_mp, const uint8_t* cur_taken, const uint8_t* mp_has_obs, float th, int32_t* cur_mp,
                                 int32_t* nmatches) {
    if (!m || !cur || !nmatches || n_mp < 0 || !scale_factors || (cur->n > 0 && !cur_mp) ||
        (n_mp > 0 && (!mp_proj || !mp_level || !mp_viewcos || !mp_desc))) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(m->device));
    *nmatches = 0;
    for (int k = 0; k < cur->n; ++k) cur_mp[k] = -1;
    if (n_mp == 0 || cur->n == 0) return FBE_OK;
    const bool bFactor = th != 1.0;
    std::vector<float4> q(n_mp);
    std::vector<int2> lv(n_mp);
    for (int i = 0; i < n_mp; ++i) {
        const int l = mp_level[i];
        if (l < 0 || l >= nlevels) return FBE_E_INVALID;
        float r = mp_viewcos[i] > 0.998 ? 2.5f : 4.0f;       // RadiusByViewingCos :132-138
        if (bFactor) r *= th;
        q[i] = make_float4(mp_proj[2 * i], mp_proj[2 * i + 1], r * scale_factors[l], 0.f);
        lv[i] = make_int2(l - 1, l);
    }
    return projection_search(m, kResolveMap, cur, q, lv, mp_desc, nullptr, n_mp, cur_taken, mp_has_obs, nullptr, 0, cur_mp, nmatches);
}

// shared body of the two SearchByBoW overloads: f_blocked[k] != 0 keeps frame / key-frame-2 feature k out of the search
// from the start (KF-KF: no good map point); strict_low selects `< TH_LOW` (:599) instead of `<= TH_LOW` (:247)
static int bow_search(fbe_matcher* m, const fbe_keypoint* kf_kps, const uint8_t* kf_desc, int32_t n_kf, const uint8_t* kf_has_mp,
                      const int32_t* kf_node_ids, const int32_t* kf_start, const int32_t* kf_items, int32_t kf_nn,
                      const fbe_keypoint* f_kps, const uint8_t* f_desc, int32_t n_f, const int32_t* f_node_ids,
                      const int32_t* f_start, const int32_t* f_items, int32_t f_nn, const uint8_t* f_blocked, int strict_low,
                      int32_t* f_mp, int32_t* nmatches) {
    if (!m || !nmatches || n_kf < 0 || n_f < 0 || kf_nn < 0 || f_nn < 0 || (n_f > 0 && !f_mp)) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(m->device));
    *nmatches = 0;
    for (int k = 0; k < n_f; ++k) f_mp[k] = -1;
    if (n_kf == 0 || n_f == 0 || kf_nn == 0 || f_nn == 0) return FBE_OK;
    if (!kf_kps || !kf_desc || !kf_has_mp || !kf_node_ids || !kf_start || !kf_items || !f_kps || !f_desc || !f_node_ids || !f_start || !f_items)
        return FBE_E_INVALID;
    // merge-walk of the two feature vectors (:184-262): the ordered query list
    std::vector<int> q_src, q_beg, q_end;
    int a = 0, b = 0, C = 1;
    while (a < kf_nn && b < f_nn) {
        if (kf_node_ids[a] == f_node_ids[b]) {
            for (int p = kf_start[a]; p < kf_start[a + 1]; ++p) {
                const int i = kf_items[p];
                if (i < 0 || i >= n_kf) return FBE_E_INVALID;
                if (!kf_has_mp[i]) continue;
                q_src.push_back(i); q_beg.push_back(f_start[b]); q_end.push_back(f_start[b + 1]);
            }
            C = std::max(C, f_start[b + 1] - f_start[b]);
            ++a; ++b;
        } else if (kf_node_ids[a] < f_node_ids[b]) ++a;
        else ++b;
    }
    const int nq = (int)q_src.size();
    if (nq == 0) return FBE_OK;
    const int n_items = f_start[f_nn];
    FBE_TRY(upload(m->fb.kps, kf_kps, (size_t)n_kf * sizeof(fbe_keypoint), m->stream));
    FBE_TRY(upload(m->fb.desc, kf_desc, (size_t)n_kf * 32, m->stream));
    FBE_TRY(upload(m->fa.kps, f_kps, (size_t)n_f * sizeof(fbe_keypoint), m->stream));
    FBE_TRY(upload(m->fa.desc, f_desc, (size_t)n_f * 32, m->stream));
    FBE_TRY(upload(m->i1, q_src.data(), (size_t)nq * 4, m->stream));
    FBE_TRY(upload(m->i2, q_beg.data(), (size_t)nq * 4, m->stream));
    FBE_TRY(upload(m->partial, q_end.data(), (size_t)nq * 4, m->stream));
    FBE_TRY(upload(m->fa.items, f_items, (size_t)n_items * 4, m->stream));
    FBE_TRY(upload(m->nq, &nq, 4, m->stream));
    FBE_TRY(upload(m->fa.n, &n_f, 4, m->stream));
    FBE_TRY(m->rows.ensure((size_t)nq * C * 4)); FBE_TRY(m->cnt.ensure((size_t)nq * 4));
    FBE_TRY(m->i0.ensure((size_t)n_f * 4)); FBE_TRY(m->i3.ensure((size_t)nq * 4)); FBE_TRY(m->i4.ensure((size_t)nq * 4));
    FBE_TRY(m->u0.ensure((size_t)n_f)); FBE_TRY(m->misc.ensure(64));
    if (f_blocked) FBE_CUDA(cudaMemcpyAsync(m->u0.p, f_blocked, (size_t)n_f, cudaMemcpyHostToDevice, m->stream));
    else FBE_CUDA(cudaMemsetAsync(m->u0.p, 0, (size_t)n_f, m->stream));
    FBE_TRY(launch_bow_rows(m->fb.desc.as<uint8_t>(), m->fa.desc.as<uint8_t>(), m->i1.as<int>(), m->i2.as<int>(), m->partial.as<int>(),
                            m->fa.items.as<int>(), nq, C, m->rows.as<unsigned>(), m->cnt.as<int>(), m->stream));
    ResolveArgs r{};
    r.mode = kResolveBow; r.C = C; r.rows = m->rows.as<unsigned>(); r.cnt = m->cnt.as<int>(); r.nq = m->nq.as<int>();
    r.q_stride = nq; r.t_stride = n_f; r.nt = m->fa.n.as<int>(); r.q_kps = m->fb.kps.as<fbe_keypoint>(); r.q_src = m->i1.as<int>();
    r.t_kps = m->fa.kps.as<fbe_keypoint>(); r.q_has_obs = nullptr; r.nn_ratio = m->nn_ratio; r.check_ori = m->check_ori;
    r.taken = m->u0.as<uint8_t>(); r.cur_mp = m->i0.as<int>(); r.q_bin = m->i3.as<int>(); r.q_hit = m->i4.as<int>();
    r.nmatches = m->misc.as<int>(); r.th_dist = strict_low ? FBE_TH_LOW - 1 : 0;
    // q_kps is indexed by the key-frame keypoint index: give the resolve kernel a q_stride-independent view
    FBE_TRY(launch_resolve(r, 1, m->stream));
    FBE_CUDA(cudaMemcpyAsync(f_mp, m->i0.p, (size_t)n_f * 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaMemcpyAsync(nmatches, m->misc.p, 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaStreamSynchronize(m->stream));
    return FBE_OK;
}

int fbe_search_by_bow(fbe_matcher* m, const fbe_keypoint* kf_kps, const uint8_t* kf_desc, int32_t n_kf, const uint8_t* kf_has_mp,
                      const int32_t* kf_node_ids, const int32_t* kf_start, const int32_t* kf_items, int32_t kf_nn,
                      const fbe_keypoint* f_kps, const uint8_t* f_desc, int32_t n_f, const int32_t* f_node_ids,
                      const int32_t* f_start, const int32_t* f_items, int32_t f_nn, int32_t* f_mp, int32_t* nmatches) {
    return bow_search(m, kf_kps, kf_desc, n_kf, kf_has_mp, kf_node_ids, kf_start, kf_items, kf_nn, f_kps, f_desc, n_f, f_node_ids, f_start,
                      f_items, f_nn, nullptr, 0, f_mp, nmatches);
}

int fbe_search_by_bow_kf(fbe_matcher* m, const fbe_keypoint* kf1_kps, const uint8_t* kf1_desc, int32_t n1, const uint8_t* kf1_has_mp,
                         const int32_t* kf1_node_ids, const int32_t* kf1_start, const int32_t* kf1_items, int32_t kf1_nn,
                         const fbe_keypoint* kf2_kps, const uint8_t* kf2_desc, int32_t n2, const uint8_t* kf2_has_mp,
                         const int32_t* kf2_node_ids, const int32_t* kf2_start, const int32_t* kf2_items, int32_t kf2_nn,
                         int32_t* matches12, int32_t* nmatches) {
    if (!m || !nmatches || n1 < 0 || n2 < 0 || (n1 > 0 && !matches12) || (n2 > 0 && !kf2_has_mp)) return FBE_E_INVALID;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<uint8_t> blocked(std::max(n2, 1));
    for (int k = 0; k < n2; ++k) blocked[k] = kf2_has_mp[k] ? 0 : 1;          // `!pMP2 || pMP2->isBad()` (:575-579)
    std::vector<int32_t> f_mp(std::max(n2, 1), -1);
    const int rc = bow_search(m, kf1_kps, kf1_desc, n1, kf1_has_mp, kf1_node_ids, kf1_start, kf1_items, kf1_nn, kf2_kps, kf2_desc, n2,
                              kf2_node_ids, kf2_start, kf2_items, kf2_nn, blocked.data(), 1, f_mp.data(), nmatches);
    if (rc != FBE_OK) return rc;
    // the search assigns every key-frame-2 feature at most once (vbMatched2) and every key-frame-1 feature at most once, so
    // the target-indexed result inverts into vpMatches12; entries removed by the orientation check stay -1 (NULL, :646)
    for (int k = 0; k < n2; ++k)
        if (f_mp[k] >= 0 && f_mp[k] < n1) matches12[f_mp[k]] = k;
    return FBE_OK;
}

int fbe_search_for_triangulation(fbe_matcher* m, const fbe_keypoint* kf1_kps, const uint8_t* kf1_desc, int32_t n1,
                                 const uint8_t* kf1_skip, const uint8_t* kf1_stereo, const int32_t* kf1_node_ids,
                                 const int32_t* kf1_start, const int32_t* kf1_items, int32_t kf1_nn, const fbe_keypoint* kf2_kps,
                                 const uint8_t* kf2_desc, int32_t n2, const uint8_t* kf2_skip, const uint8_t* kf2_stereo,
                                 const int32_t* kf2_node_ids, const int32_t* kf2_start, const int32_t* kf2_items, int32_t kf2_nn,
                                 const float F12[9], float ex, float ey, const float* kf2_scale_factors,
                                 const float* kf2_level_sigma2, int32_t nlevels, int32_t* matches12, int32_t* nmatches) {
    if (!m || !nmatches || n1 < 0 || n2 < 0 || kf1_nn < 0 || kf2_nn < 0 || (n1 > 0 && !matches12) || !F12 || !kf2_scale_factors ||
        !kf2_level_sigma2 || nlevels < 1 || nlevels > FBE_MAX_LEVELS)
        return FBE_E_INVALID;
    *nmatches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0 || kf1_nn == 0 || kf2_nn == 0) return FBE_OK;
    if (!kf1_kps || !kf1_desc || !kf1_skip || !kf1_stereo || !kf1_node_ids || !kf1_start || !kf1_items || !kf2_kps || !kf2_desc ||
        !kf2_skip || !kf2_stereo || !kf2_node_ids || !kf2_start || !kf2_items)
        return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(m->device));
    for (int k = 0; k < n2; ++k)
        if (kf2_kps[k].octave < 0 || kf2_kps[k].octave >= nlevels) { set_error("key frame 2 keypoint octave outside the scale tables"); return FBE_E_INVALID; }
    // merge-walk of the two feature vectors (:694-776; lower_bound jumps land where ++ would): the query list
    std::vector<int> q_src, q_beg, q_end;
    int a = 0, b = 0;
    while (a < kf1_nn && b < kf2_nn) {
        if (kf1_node_ids[a] == kf2_node_ids[b]) {
            if (kf2_start[b + 1] - kf2_start[b] >= (1 << 20)) { set_error("more than 2^20 features under one vocabulary node"); return FBE_E_UNSUPPORTED; }
            for (int p = kf1_start[a]; p < kf1_start[a + 1]; ++p) {
                const int i = kf1_items[p];
                if (i < 0 || i >= n1) return FBE_E_INVALID;
                if (kf1_skip[i]) continue;
                q_src.push_back(i); q_beg.push_back(kf2_start[b]); q_end.push_back(kf2_start[b + 1]);
            }
            ++a; ++b;
        } else if (kf1_node_ids[a] < kf2_node_ids[b]) ++a;
        else ++b;
    }
    const int nq = (int)q_src.size();
    if (nq == 0) return FBE_OK;
    const int n_items = kf2_start[kf2_nn];
    for (int p = 0; p < n_items; ++p)
        if (kf2_items[p] < 0 || kf2_items[p] >= n2) return FBE_E_INVALID;
    FBE_TRY(upload(m->fb.kps, kf1_kps, (size_t)n1 * sizeof(fbe_keypoint), m->stream));
    FBE_TRY(upload(m->fb.desc, kf1_desc, (size_t)n1 * 32, m->stream));
    FBE_TRY(upload(m->fa.kps, kf2_kps, (size_t)n2 * sizeof(fbe_keypoint), m->stream));
    FBE_TRY(upload(m->fa.desc, kf2_desc, (size_t)n2 * 32, m->stream));
    FBE_TRY(upload(m->fa.items, kf2_items, (size_t)n_items * 4, m->stream));
    FBE_TRY(upload(m->i1, q_src.data(), (size_t)nq * 4, m->stream));
    FBE_TRY(upload(m->i2, q_beg.data(), (size_t)nq * 4, m->stream));
    FBE_TRY(upload(m->partial, q_end.data(), (size_t)nq * 4, m->stream));
    FBE_TRY(upload(m->u0, kf2_skip, (size_t)n2, m->stream));
    // the two stereo flag arrays share one buffer: [kf1 | kf2]
    FBE_TRY(m->u1.ensure((size_t)n1 + (size_t)n2 + 16));
    FBE_CUDA(cudaMemcpyAsync(m->u1.p, kf1_stereo, (size_t)n1, cudaMemcpyHostToDevice, m->stream));
    FBE_CUDA(cudaMemcpyAsync(m->u1.as<uint8_t>() + n1, kf2_stereo, (size_t)n2, cudaMemcpyHostToDevice, m->stream));
    FBE_TRY(m->i0.ensure((size_t)n1 * 4)); FBE_TRY(m->i3.ensure((size_t)nq * 4)); FBE_TRY(m->i4.ensure((size_t)nq * 4));
    FBE_TRY(m->misc.ensure(64));
    FBE_CUDA(cudaMemsetAsync(m->i0.p, 0xFF, (size_t)n1 * 4, m->stream));
    FBE_CUDA(cudaMemsetAsync(m->misc.p, 0, 4, m->stream));
    TriArgs t{};
    t.kps1 = m->fb.kps.as<fbe_keypoint>(); t.desc1 = m->fb.desc.as<uint8_t>(); t.stereo1 = m->u1.as<uint8_t>();
    t.kps2 = m->fa.kps.as<fbe_keypoint>(); t.desc2 = m->fa.desc.as<uint8_t>(); t.stereo2 = m->u1.as<uint8_t>() + n1;
    t.skip2 = m->u0.as<uint8_t>(); t.items2 = m->fa.items.as<int>();
    t.q_src = m->i1.as<int>(); t.q_beg = m->i2.as<int>(); t.q_end = m->partial.as<int>(); t.nq = nq;
    for (int k = 0; k < 9; ++k) t.F[k] = F12[k];
    t.ex = ex; t.ey = ey;
    for (int k = 0; k < FBE_MAX_LEVELS; ++k) { t.scale[k] = k < nlevels ? kf2_scale_factors[k] : 0.f; t.sigma2[k] = k < nlevels ? kf2_level_sigma2[k] : 0.f; }
    t.check_ori = m->check_ori;
    t.q_best = m->i3.as<int>(); t.q_bin = m->i4.as<int>(); t.matches12 = m->i0.as<int>(); t.nmatches = m->misc.as<int>();
    FBE_TRY(launch_triangulation(t, m->stream));
    FBE_CUDA(cudaMemcpyAsync(matches12, m->i0.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaMemcpyAsync(nmatches, m->misc.p, 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaStreamSynchronize(m->stream));
    return FBE_OK;
}

int fbe_fuse_search(fbe_matcher* m, const fbe_frame_view* kf, const float* kf_uright, const float* inv_level_sigma2, int32_t nlevels,
                    const float* proj, const float* proj_ur, const int32_t* level, const float* radius, const uint8_t* mp_desc,
                    int32_t n, int32_t check_chi2, int32_t* best_idx, int32_t* best_dist) {
    if (!m || !kf || n < 0 || (n > 0 && (!proj || !level || !radius || !mp_desc || !best_idx || !best_dist)) ||
        (check_chi2 && (!inv_level_sigma2 || nlevels < 1 || nlevels > FBE_MAX_LEVELS)))
        return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(m->device));
    for (int i = 0; i < n; ++i) { best_idx[i] = -1; best_dist[i] = INT_MAX; }
    if (n == 0 || kf->n == 0) return FBE_OK;
    if (check_chi2)
        for (int k = 0; k < kf->n; ++k)
            if (kf->kps[k].octave < 0 || kf->kps[k].octave >= nlevels) { set_error("key frame keypoint octave outside mvInvLevelSigma2"); return FBE_E_INVALID; }
    std::vector<float4> q(n);
    std::vector<int2> lv(n);
    for (int i = 0; i < n; ++i) {
        const bool skip = std::isnan(proj[2 * i]);
        q[i] = make_float4(skip ? 0.f : proj[2 * i], skip ? 0.f : proj[2 * i + 1], skip ? -1.f : radius[i], proj_ur ? proj_ur[i] : 0.f);
        // `kpLevel<nPredictedLevel-1 || kpLevel>nPredictedLevel` (:905, :1065); level 0 -> minLevel -1 filters nothing below
        lv[i] = make_int2(level[i] - 1, level[i]);
    }
    FrameDev Kf;
    QueryDev qs;
    FBE_TRY(upload_frame(m, m->fa, kf, true, Kf));
    FBE_TRY(upload_queries(m, q, lv, mp_desc, n, qs));
    FBE_TRY(m->i0.ensure((size_t)n * 4)); FBE_TRY(m->i1.ensure((size_t)n * 4)); FBE_TRY(m->i2.ensure((size_t)n * 4));
    if (check_chi2) {
        ReprojGate rg;
        for (int k = 0; k < nlevels; ++k) rg.inv_sigma2[k] = inv_level_sigma2[k];
        if (kf_uright) { FBE_TRY(upload(m->f0, kf_uright, (size_t)kf->n * 4, m->stream)); rg.t_uright = m->f0.as<float>(); }
        FBE_TRY(launch_window_top2_reproj(Kf, qs, n, rg, m->i0.as<int>(), m->i1.as<int>(), m->i2.as<int>(), m->stream));
    } else {
        FBE_TRY(launch_window_top2(Kf, qs, 1, n, true, m->i0.as<int>(), m->i1.as<int>(), m->i2.as<int>(), m->stream));
    }
    FBE_CUDA(cudaMemcpyAsync(best_idx, m->i0.p, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaMemcpyAsync(best_dist, m->i1.p, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaStreamSynchronize(m->stream));
    return FBE_OK;
}

int fbe_distinctive_descriptors(fbe_matcher* m, const uint8_t* desc, const int32_t* start, int32_t npts, int32_t* best,
                                int32_t* best_median) {
    if (!m || npts < 0 || (npts > 0 && (!start || !best))) return FBE_E_INVALID;
    if (npts == 0) return FBE_OK;
    const int total = start[npts];
    if (start[0] != 0 || total < 0 || (total > 0 && !desc)) return FBE_E_INVALID;
    for (int p = 0; p < npts; ++p) {
        if (start[p + 1] < start[p]) return FBE_E_INVALID;
        if (start[p + 1] - start[p] >= (1 << 20)) { set_error("more than 2^20 observations of one map point"); return FBE_E_UNSUPPORTED; }
    }
    FBE_CUDA(cudaSetDevice(m->device));
    FBE_TRY(upload(m->fa.desc, desc, (size_t)total * 32, m->stream));
    FBE_TRY(upload(m->i1, start, (size_t)(npts + 1) * 4, m->stream));
    FBE_TRY(m->i2.ensure((size_t)npts * 4)); FBE_TRY(m->i3.ensure((size_t)npts * 4));
    FBE_TRY(launch_distinctive(m->fa.desc.as<uint8_t>(), m->i1.as<int>(), npts, m->i2.as<int>(), m->i3.as<int>(), m->stream));
    FBE_CUDA(cudaMemcpyAsync(best, m->i2.p, (size_t)npts * 4, cudaMemcpyDeviceToHost, m->stream));
    if (best_median) FBE_CUDA(cudaMemcpyAsync(best_median, m->i3.p, (size_t)npts * 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaStreamSynchronize(m->stream));
    return FBE_OK;
}

int fbe_bruteforce_top2(fbe_matcher* m, const uint8_t* q_desc, int32_t nq, const uint8_t* t_desc, int32_t nt, int32_t* best_idx,
                        int32_t* best_dist, int32_t* second_dist) {
    if (!m || nq < 0 || nt < 0 || (nq > 0 && (!q_desc || !best_idx || !best_dist || !second_dist)) || (nt > 0 && !t_desc)) return FBE_E_INVALID;
    if (nt >= (1 << 20)) { set_error("more than 2^20 targets"); return FBE_E_UNSUPPORTED; }
    FBE_CUDA(cudaSetDevice(m->device));
    if (nq == 0) return FBE_OK;
    const int qblocks = (nq + 127) / 128;
    // enough CTAs for ~32 resident warps per SM: the POPC pipe (16 lanes/clk/SM) needs many warps in flight to stay busy
    int nchunks = std::max(1, std::min((8 * 148 + qblocks - 1) / qblocks, (nt + 255) / 256));
    FBE_TRY(upload(m->qdesc, q_desc, (size_t)nq * 32, m->stream));
    FBE_TRY(upload(m->fa.desc, t_desc, (size_t)nt * 32, m->stream));
    FBE_TRY(m->partial.ensure((size_t)nchunks * nq * 8));
    FBE_TRY(m->i0.ensure((size_t)nq * 4)); FBE_TRY(m->i1.ensure((size_t)nq * 4)); FBE_TRY(m->i2.ensure((size_t)nq * 4));
    FBE_TRY(launch_bruteforce(m->qdesc.as<uint8_t>(), nq, m->fa.desc.as<uint8_t>(), nt, m->partial.as<unsigned>(), nchunks,
                              m->i0.as<int>(), m->i1.as<int>(), m->i2.as<int>(), m->stream));
    FBE_CUDA(cudaMemcpyAsync(best_idx, m->i0.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaMemcpyAsync(best_dist, m->i1.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaMemcpyAsync(second_dist, m->i2.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, m->stream));
    FBE_CUDA(cudaStreamSynchronize(m->stream));
    return FBE_OK;
}

}  // extern "C"
