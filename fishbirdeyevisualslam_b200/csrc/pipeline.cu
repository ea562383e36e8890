// Device-resident batch pipeline (config C2/C4): per step `batch` front+bird frame pairs are extracted, bucketed and
// matched against the previous pair without any host round trip.
//   front : ORBextractor -> 64x48 grid -> SearchForInitialization-style window search (vbPrevMatched = previous keypoints)
//   bird  : ORBextractor -> 32x32 grid -> BirdviewMatch (isProject == 0)
// Frame slots: each extractor workspace holds batch+1 frames; slot 0 carries the last frame of the previous step, slots
// 1..batch receive this step's frames, so "pair p vs pair p-1" is the uniform-stride problem (query slot p, target
// slot p+1) for p = 0..batch-1.  The bird extractor runs on its own stream beside the front extractor.
// Matching runs on a third stream and the extractor OUTPUT arrays exist twice (sets alternate per step), so the
// latency-bound tail of step N (window rows, sequential resolve, bird top-2) overlaps the extraction of step N+1.
#include <cstring>
#include <new>
#include <string>
#include "match_kernels.cuh"

using namespace fbe;

struct fbe_pipeline {
    fbe_pipeline_cfg cfg;
    ExtractorCore front, bird;
    int B = 0, fcap = 0, bcap = 0, row_cap = 256;
    bool have_prev = false;
    cudaStream_t mstream = nullptr;      // matching
    cudaEvent_t ev_front = nullptr, ev_bird = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_match_done[2] = {nullptr, nullptr};   // per output set: the matching that read it has finished
    bool set_used[2] = {false, false};
    int step_count = 0, last_set = 0;
    float bounds[4] = {0, 0, 0, 0};      // mnMinX, mnMaxX, mnMinY, mnMaxY of the front frames
    // front matching workspace
    float4* fq = nullptr; int2* flv = nullptr; unsigned* frows = nullptr; int* fcnt = nullptr;
    int *f_mdist = nullptr, *f_m21 = nullptr, *f_m12 = nullptr, *f_bin = nullptr, *f_hit = nullptr, *f_nm = nullptr;
    // bird matching workspace
    float4* bq = nullptr; int2* blv = nullptr;
    int *b_bi = nullptr, *b_bd = nullptr, *b_sd = nullptr, *b_m12 = nullptr, *b_bin = nullptr, *b_nm = nullptr;
    int* flags = nullptr;                // [0] row overflow
    fbe_pair_result* d_res = nullptr;
    uint8_t *d_front_in = nullptr, *d_bird_in = nullptr;   // staging for the host-buffer step
    fbe_pair_result* h_res = nullptr;    // pinned
    // asynchronous host-buffer path: the H2D copy of step N+1 runs on its own stream beside the kernels of step N
    cudaStream_t copy_stream = nullptr;
    cudaStream_t out_stream = nullptr;   // D2H of the extractor outputs (keypoints + descriptors) beside the matching
    cudaEvent_t ev_feat_done[2] = {nullptr, nullptr};    // per output set: its feature copy-out has finished
    bool feat_used[2] = {false, false};
    int step_seq = 0;                    // steps submitted so far (error reports)
    static constexpr int kQ = 3;         // steps in flight on the host-buffer path (input staging ring)
    uint8_t *d_front_q[kQ] = {}, *d_bird_q[kQ] = {};
    cudaEvent_t ev_in_ready[kQ] = {}, ev_in_free[kQ] = {}, ev_done[kQ] = {};
    cudaEvent_t ev_copy0[kQ] = {};       // timing: start of the input copy (ev_in_ready is its end)
    cudaEvent_t ev_feat_q[kQ] = {};      // per ticket: the feature copy-out of that step has finished
    bool feat_q_used[kQ] = {};
    bool in_used[kQ] = {};
    int* h_flags = nullptr;              // pinned, [kQ][2]
    int32_t next_ticket = 0;
};

namespace {

#define FBE_TRY(expr) do { int _rc = (expr); if (_rc != FBE_OK) return _rc; } while (0)

__global__ void k_pair_results(const int* __restrict__ nf, const int* __restrict__ nb, const int* __restrict__ fm,
                               const int* __restrict__ bm, const int* __restrict__ fstatus, const int* __restrict__ bstatus,
                               int B, fbe_pair_result* __restrict__ out, int* __restrict__ flags) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    fbe_pair_result r;
    r.n_front = nf[p + 1]; r.n_bird = nb[p + 1]; r.front_matches = fm[p]; r.bird_matches = bm[p];
    out[p] = r;
    // octree workspace overflow of either extractor (slot p + 1 of this step): flags[2] = 1 + pair
    if (fstatus[p] | bstatus[p]) atomicMax(flags + 2, p + 1);
}

// the per-step flag record copied to the host with every result: [0] candidate-row overflow, [2] 1 + pair whose octree
// workspace overflowed, [4] 1 + (pair << 16 | query) of one overflowing window
constexpr int kFlagInts = 8;
int check_flags(const int* h, int row_cap) {
    if (h[0]) {
        const int w = h[4] - 1;
        set_error("candidate rows overflow in pipeline step: pair " + std::to_string(w >> 16) + ", front query " + std::to_string(w & 0xFFFF) +
                  " has more than " + std::to_string(row_cap) + " keypoints in its search window (raise fbe_pipeline_cfg.front_row_cap)");
        return FBE_E_CAPACITY;
    }
    if (h[2]) { set_error("octree workspace overflow in pipeline step: pair " + std::to_string(h[2] - 1)); return FBE_E_CAPACITY; }
    return FBE_OK;
}

FrameDev frame_dev(const ExtractorCore& e, int slot0, int set) {
    Workspace v = e.slot_view(slot0, set);
    FrameDev f;
    f.kps = v.out_kps_un; f.desc = v.out_desc; f.n = v.out_n; f.start = v.grid_start; f.items = v.grid_items;     // matchers see mvKeysUn
    f.kp_stride = e.hplan.kp_cap_total;
    f.min_x = e.hplan.grid_min_x; f.min_y = e.hplan.grid_min_y; f.inv_w = e.hplan.grid_inv_w; f.inv_h = e.hplan.grid_inv_h;
    f.gcols = e.hplan.grid_cols; f.grows = e.hplan.grid_rows;
    return f;
}

// last frame of this step (slot B of `set`) becomes slot 0 of the other set, where the next step's matching reads it
int carry_last(ExtractorCore& e, int B, int set, cudaStream_t st) {
    Workspace s = e.slot_view(B, set), d = e.slot_view(0, 1 - set);
    const int cap = e.hplan.kp_cap_total, gcells = e.hplan.grid_cols * e.hplan.grid_rows;
    FBE_CUDA(cudaMemcpyAsync(d.out_kps, s.out_kps, (size_t)cap * sizeof(fbe_keypoint), cudaMemcpyDeviceToDevice, st));
    if (s.out_kps_un != s.out_kps) FBE_CUDA(cudaMemcpyAsync(d.out_kps_un, s.out_kps_un, (size_t)cap * sizeof(fbe_keypoint), cudaMemcpyDeviceToDevice, st));
    FBE_CUDA(cudaMemcpyAsync(d.out_desc, s.out_desc, (size_t)cap * 32, cudaMemcpyDeviceToDevice, st));
    FBE_CUDA(cudaMemcpyAsync(d.out_n, s.out_n, sizeof(int), cudaMemcpyDeviceToDevice, st));
    FBE_CUDA(cudaMemcpyAsync(d.grid_start, s.grid_start, (size_t)(gcells + 1) * sizeof(int), cudaMemcpyDeviceToDevice, st));
    FBE_CUDA(cudaMemcpyAsync(d.grid_items, s.grid_items, (size_t)cap * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return FBE_OK;
}

void free_all(fbe_pipeline* p) {
    cudaSetDevice(p->cfg.device);
    if (p->front.stream) cudaStreamSynchronize(p->front.stream);
    if (p->bird.stream) cudaStreamSynchronize(p->bird.stream);
    if (p->mstream) { cudaStreamSynchronize(p->mstream); cudaStreamDestroy(p->mstream); p->mstream = nullptr; }
    void* ptrs[] = {p->fq, p->flv, p->frows, p->fcnt, p->f_mdist, p->f_m21, p->f_m12, p->f_bin, p->f_hit, p->f_nm, p->bq, p->blv,
                    p->b_bi, p->b_bd, p->b_sd, p->b_m12, p->b_bin, p->b_nm, p->flags, p->d_res, p->d_front_in, p->d_bird_in};
    for (void* q : ptrs) if (q) cudaFree(q);
    if (p->h_res) cudaFreeHost(p->h_res);
    if (p->h_flags) cudaFreeHost(p->h_flags);
    if (p->copy_stream) { cudaStreamSynchronize(p->copy_stream); cudaStreamDestroy(p->copy_stream); }
    if (p->out_stream) { cudaStreamSynchronize(p->out_stream); cudaStreamDestroy(p->out_stream); }
    for (cudaEvent_t e : {p->ev_feat_done[0], p->ev_feat_done[1]}) if (e) cudaEventDestroy(e);
    for (int k = 0; k < fbe_pipeline::kQ; ++k) {
        if (p->d_front_q[k]) cudaFree(p->d_front_q[k]);
        if (p->d_bird_q[k]) cudaFree(p->d_bird_q[k]);
        for (cudaEvent_t e : {p->ev_in_ready[k], p->ev_in_free[k], p->ev_done[k], p->ev_copy0[k], p->ev_feat_q[k]}) if (e) cudaEventDestroy(e);
    }
    for (cudaEvent_t e : {p->ev_front, p->ev_bird, p->ev_match_done[0], p->ev_match_done[1], p->ev_t0, p->ev_t1}) if (e) cudaEventDestroy(e);
    p->front.destroy();
    p->bird.destroy();
}

}  // namespace

extern "C" {

int fbe_pipeline_create(const fbe_pipeline_cfg* cfg, fbe_pipeline** out) {
    if (!cfg || !out || cfg->batch < 1) return FBE_E_INVALID;
    *out = nullptr;
    fbe_pipeline* p = new (std::nothrow) fbe_pipeline();
    if (!p) return FBE_E_INVALID;
    p->cfg = *cfg;
    p->B = cfg->batch;
    if (cfg->front_row_cap > 0) p->row_cap = cfg->front_row_cap;
    fbe_extractor_cfg fc = cfg->front, bc = cfg->bird;
    fc.max_batch = bc.max_batch = cfg->batch + 1;
    fc.device = bc.device = cfg->device;
    p->front.out_sets = p->bird.out_sets = 2;
    // The front extractor's main stream (pyramid -> FAST -> octree -> describe -> grid) is the critical path of a step: the
    // blur side stream, the bird extractor and the matching fill in around it.  FBE_STREAM_PRIO=0 disables (A/B runs).
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        const char* pe = getenv("FBE_STREAM_PRIO");
        if (!(pe && pe[0] == '0')) p->front.stream_priority = prio_hi;
    }
    int rc = p->front.init(fc);
    if (rc == FBE_OK) rc = p->bird.init(bc);
    auto fail = [&](int code) { free_all(p); delete p; return code; };
    if (rc != FBE_OK) return fail(rc);
    // k1 == 0 image bounds: mnMinX = 0, mnMaxX = cols (src/Frame.cc:741-795); grid constants of src/Frame.cc:276-283
    p->bounds[0] = 0.f; p->bounds[1] = (float)cfg->front_cols; p->bounds[2] = 0.f; p->bounds[3] = (float)cfg->front_rows;
    if (cfg->front_fisheye) {
        if ((rc = p->front.set_fisheye(cfg->front_K, cfg->front_D)) != FBE_OK) return fail(rc);
        if ((rc = fbe_image_bounds(cfg->front_cols, cfg->front_rows, cfg->front_K, cfg->front_D, cfg->device, p->bounds)) != FBE_OK) return fail(rc);
    }
    // mfGridElementWidthInv = FRAME_GRID_COLS / (mnMaxX - mnMinX), src/Frame.cc:276-283
    if ((rc = p->front.set_grid(p->bounds[0], p->bounds[2], 64.f / (p->bounds[1] - p->bounds[0]), 48.f / (p->bounds[3] - p->bounds[2]), 64, 48)) != FBE_OK) return fail(rc);
    if ((rc = p->bird.set_grid(0.f, 0.f, 32.f / (float)cfg->bird_cols, 32.f / (float)cfg->bird_rows, 32, 32)) != FBE_OK) return fail(rc);
    if ((rc = p->front.ensure_plan(cfg->front_rows, cfg->front_cols)) != FBE_OK) return fail(rc);
    if ((rc = p->bird.ensure_plan(cfg->bird_rows, cfg->bird_cols)) != FBE_OK) return fail(rc);
    p->fcap = p->front.hplan.kp_cap_total;
    p->bcap = p->bird.hplan.kp_cap_total;
    const size_t B = (size_t)p->B, fc_ = (size_t)p->fcap, bc_ = (size_t)p->bcap;
    bool ok = true;
    auto A = [&](auto** ptr, size_t bytes) { if (ok && cudaMalloc((void**)ptr, bytes) != cudaSuccess) ok = false; };
    A(&p->fq, (B + 1) * fc_ * sizeof(float4)); A(&p->flv, (B + 1) * fc_ * sizeof(int2));
    A(&p->frows, (B + 1) * fc_ * p->row_cap * 4); A(&p->fcnt, (B + 1) * fc_ * 4);
    A(&p->f_mdist, (B + 1) * fc_ * 4); A(&p->f_m21, (B + 1) * fc_ * 4); A(&p->f_m12, (B + 1) * fc_ * 4);
    A(&p->f_bin, (B + 1) * fc_ * 4); A(&p->f_hit, (B + 1) * fc_ * 4); A(&p->f_nm, (B + 1) * 4);
    A(&p->bq, (B + 1) * bc_ * sizeof(float4)); A(&p->blv, (B + 1) * bc_ * sizeof(int2));
    A(&p->b_bi, (B + 1) * bc_ * 4); A(&p->b_bd, (B + 1) * bc_ * 4); A(&p->b_sd, (B + 1) * bc_ * 4);
    A(&p->b_m12, (B + 1) * bc_ * 4); A(&p->b_bin, (B + 1) * bc_ * 4); A(&p->b_nm, (B + 1) * 4);
    A(&p->flags, 64); A(&p->d_res, B * sizeof(fbe_pair_result));
    if (!ok) { set_error("pipeline workspace allocation failed"); return fail(FBE_E_CUDA); }
    if (cudaMallocHost((void**)&p->h_res, B * sizeof(fbe_pair_result)) != cudaSuccess) { set_error("pinned alloc failed"); return fail(FBE_E_CUDA); }
    cudaMemset(p->flags, 0, 64);
    cudaMemset(p->front.ws.out_n, 0, 2 * (B + 1) * sizeof(int));
    cudaMemset(p->bird.ws.out_n, 0, 2 * (B + 1) * sizeof(int));
    cudaMemset(p->front.ws.grid_start, 0, 2 * (B + 1) * (64 * 48 + 1) * sizeof(int));
    cudaMemset(p->bird.ws.grid_start, 0, 2 * (B + 1) * (32 * 32 + 1) * sizeof(int));
    if (cudaStreamCreateWithFlags(&p->mstream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); return fail(FBE_E_CUDA); }
    for (cudaEvent_t* e : {&p->ev_front, &p->ev_bird, &p->ev_match_done[0], &p->ev_match_done[1]}) cudaEventCreateWithFlags(e, cudaEventDisableTiming);
    cudaEventCreate(&p->ev_t0); cudaEventCreate(&p->ev_t1);
    cudaDeviceSynchronize();
    *out = p;
    return FBE_OK;
}

int fbe_pipeline_destroy(fbe_pipeline* p) {
    if (!p) return FBE_E_INVALID;
    free_all(p);
    delete p;
    return FBE_OK;
}

int fbe_pipeline_caps(const fbe_pipeline* p, int32_t* front_cap, int32_t* bird_cap) {
    if (!p) return FBE_E_INVALID;
    if (front_cap) *front_cap = p->fcap;
    if (bird_cap) *bird_cap = p->bcap;
    return FBE_OK;
}

int fbe_pipeline_stream(fbe_pipeline* p, void** stream) {
    if (!p || !stream) return FBE_E_INVALID;
    *stream = (void*)p->front.stream;
    return FBE_OK;
}

int fbe_pipeline_step_dev(fbe_pipeline* p, const uint8_t* d_front, const uint8_t* d_bird) {
    if (!p || !d_front || !d_bird) return FBE_E_INVALID;
    const fbe_pipeline_cfg& c = p->cfg;
    FBE_CUDA(cudaSetDevice(c.device));
    cudaStream_t fs = p->front.stream, bs = p->bird.stream, ms = p->mstream;
    const int B = p->B;
    const int set = p->step_count & 1;
    FBE_CUDA(cudaEventRecord(p->ev_t0, fs));
    // the extractors must not overwrite an output set that the matching of two steps ago (same set) still reads
    if (p->set_used[set]) {
        FBE_CUDA(cudaStreamWaitEvent(fs, p->ev_match_done[set], 0));
        FBE_CUDA(cudaStreamWaitEvent(bs, p->ev_match_done[set], 0));
    }
    if (p->feat_used[set]) {             // ... nor one whose copy-out to the host is still running
        FBE_CUDA(cudaStreamWaitEvent(fs, p->ev_feat_done[set], 0));
        FBE_CUDA(cudaStreamWaitEvent(bs, p->ev_feat_done[set], 0));
        p->feat_used[set] = false;
    }
    FBE_TRY(p->front.run_dev(d_front, c.front_cols, c.front_rows * c.front_cols, B, c.front_rows, c.front_cols, 1, set));
    FBE_CUDA(cudaEventRecord(p->ev_front, fs));
    FBE_TRY(p->bird.run_dev(d_bird, c.bird_cols, c.bird_rows * c.bird_cols, B, c.bird_rows, c.bird_cols, 1, set));
    FBE_CUDA(cudaEventRecord(p->ev_bird, bs));

    // ---- front: pair p (slot p+1) against pair p-1 (slot p) --------------------------------------------------------
    FBE_CUDA(cudaStreamWaitEvent(ms, p->ev_front, 0));
    // flags of THIS step only ([0] rows overflow, [1] first offending query, [2] octree workspace overflow): every fetch /
    // wait reads them after the step's last kernel on this stream, so one dense frame does not poison the handle
    FBE_CUDA(cudaMemsetAsync(p->flags, 0, 64, ms));
    const FrameDev fq = frame_dev(p->front, 0, set), ft = frame_dev(p->front, 1, set);
    FBE_TRY(launch_queries_from_kps(fq.kps, nullptr, nullptr, fq.n, p->fcap, B, (float)c.front_window, p->fq, p->flv, ms));
    QueryDev fqs{p->fq, p->flv, fq.desc, fq.n, p->fcap};
    FBE_TRY(launch_window_rows(ft, fqs, B, p->fcap, true, p->row_cap, p->frows, p->fcnt, p->flags, ms));
    ResolveArgs a{};
    a.mode = kResolveInit; a.C = p->row_cap; a.rows = p->frows; a.cnt = p->fcnt; a.nq = fq.n; a.q_stride = p->fcap;
    a.t_stride = p->fcap; a.nt = ft.n; a.q_kps = fq.kps; a.t_kps = ft.kps; a.nn_ratio = c.nn_ratio; a.check_ori = c.check_orientation;
    a.matched_dist = p->f_mdist; a.match21 = p->f_m21; a.matches12 = p->f_m12; a.prev_matched = nullptr;
    a.q_bin = p->f_bin; a.q_hit = p->f_hit; a.nmatches = p->f_nm;
    FBE_TRY(launch_resolve(a, B, ms));

    // ---- bird ----------------------------------------------------------------------------------------------------
    FBE_CUDA(cudaStreamWaitEvent(ms, p->ev_bird, 0));
    const FrameDev bq = frame_dev(p->bird, 0, set), bt = frame_dev(p->bird, 1, set);
    FBE_TRY(launch_queries_from_kps(bq.kps, nullptr, nullptr, bq.n, p->bcap, B, (float)c.bird_window, p->bq, p->blv, ms));
    QueryDev bqs{p->bq, p->blv, bq.desc, bq.n, p->bcap};
    FBE_TRY(launch_window_top2(bt, bqs, B, p->bcap, false, p->b_bi, p->b_bd, p->b_sd, ms));
    BirdFinishArgs f{};
    f.best_idx = p->b_bi; f.best_dist = p->b_bd; f.second_dist = p->b_sd; f.nq = bq.n; f.q_stride = p->bcap;
    f.q_kps = bq.kps; f.t_kps = bt.kps; f.t_stride = p->bcap; f.nn_ratio = c.nn_ratio; f.check_ori = c.check_orientation;
    f.matches12 = p->b_m12; f.dmatches = nullptr; f.n_dmatches = nullptr; f.nmatches = p->b_nm; f.q_bin = p->b_bin;
    FBE_TRY(launch_bird_finish(f, B, ms));

    k_pair_results<<<(B + 127) / 128, 128, 0, ms>>>(fq.n, bq.n, p->f_nm, p->b_nm, p->front.slot_view(1, set).status,
                                                   p->bird.slot_view(1, set).status, B, p->d_res, p->flags);
    count_launch();
    FBE_CUDA(cudaGetLastError());
    FBE_TRY(carry_last(p->front, B, set, ms));
    FBE_TRY(carry_last(p->bird, B, set, ms));
    FBE_CUDA(cudaEventRecord(p->ev_match_done[set], ms));
    FBE_CUDA(cudaEventRecord(p->ev_t1, ms));
    p->set_used[set] = true;
    p->last_set = set;
    p->step_count++;
    p->have_prev = true;
    return FBE_OK;
}

// Orders the front stream (the one fbe_pipeline_stream() hands out) after everything enqueued so far, without a host
// synchronisation, so that a caller's event recorded on it afterwards marks the end of the submitted steps.
int fbe_pipeline_join(fbe_pipeline* p) {
    if (!p) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(p->cfg.device));
    if (p->step_count > 0) FBE_CUDA(cudaStreamWaitEvent(p->front.stream, p->ev_match_done[p->last_set], 0));
    return FBE_OK;
}

int fbe_pipeline_sync(fbe_pipeline* p) {
    if (!p) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(p->cfg.device));
    FBE_CUDA(cudaStreamSynchronize(p->front.stream));
    FBE_CUDA(cudaStreamSynchronize(p->bird.stream));
    FBE_CUDA(cudaStreamSynchronize(p->mstream));
    return FBE_OK;
}

int fbe_pipeline_last_step_ms(fbe_pipeline* p, float* ms) {
    if (!p || !ms) return FBE_E_INVALID;
    FBE_CUDA(cudaEventSynchronize(p->ev_t1));
    FBE_CUDA(cudaEventElapsedTime(ms, p->ev_t0, p->ev_t1));
    return FBE_OK;
}

int fbe_pipeline_fetch(fbe_pipeline* p, fbe_pair_result* res, int32_t* front_matches12, int32_t* bird_matches12) {
    if (!p) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(p->cfg.device));
    cudaStream_t ms = p->mstream;
    const size_t B = (size_t)p->B;
    int h_flags[kFlagInts] = {};
    FBE_CUDA(cudaMemcpyAsync(p->h_res, p->d_res, B * sizeof(fbe_pair_result), cudaMemcpyDeviceToHost, ms));
    FBE_CUDA(cudaMemcpyAsync(h_flags, p->flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ms));
    if (front_matches12) FBE_CUDA(cudaMemcpyAsync(front_matches12, p->f_m12, B * p->fcap * 4, cudaMemcpyDeviceToHost, ms));
    if (bird_matches12) FBE_CUDA(cudaMemcpyAsync(bird_matches12, p->b_m12, B * p->bcap * 4, cudaMemcpyDeviceToHost, ms));
    FBE_CUDA(cudaStreamSynchronize(ms));
    FBE_TRY(check_flags(h_flags, p->row_cap));
    if (res) std::memcpy(res, p->h_res, B * sizeof(fbe_pair_result));
    return FBE_OK;
}

int fbe_pipeline_step_host(fbe_pipeline* p, const uint8_t* h_front, const uint8_t* h_bird, fbe_pair_result* res,
                           int32_t* front_matches12, int32_t* bird_matches12) {
    if (!p || !h_front || !h_bird) return FBE_E_INVALID;
    const fbe_pipeline_cfg& c = p->cfg;
    FBE_CUDA(cudaSetDevice(c.device));
    const size_t fbytes = (size_t)p->B * c.front_rows * c.front_cols, bbytes = (size_t)p->B * c.bird_rows * c.bird_cols;
    if (!p->d_front_in) FBE_CUDA(cudaMalloc(&p->d_front_in, fbytes));
    if (!p->d_bird_in) FBE_CUDA(cudaMalloc(&p->d_bird_in, bbytes));
    // the previous step may still be reading the staging buffers
    FBE_TRY(fbe_pipeline_sync(p));
    FBE_CUDA(cudaMemcpyAsync(p->d_front_in, h_front, fbytes, cudaMemcpyHostToDevice, p->front.stream));
    FBE_CUDA(cudaMemcpyAsync(p->d_bird_in, h_bird, bbytes, cudaMemcpyHostToDevice, p->bird.stream));
    FBE_TRY(fbe_pipeline_step_dev(p, p->d_front_in, p->d_bird_in));
    return fbe_pipeline_fetch(p, res, front_matches12, bird_matches12);
}

int fbe_pipeline_fetch_front_undistorted(fbe_pipeline* p, int32_t pair, fbe_keypoint* front_kps_un, float bounds[4]) {
    if (!p || pair < 0 || pair >= p->B) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(p->cfg.device));
    FBE_TRY(fbe_pipeline_sync(p));
    Workspace f = p->front.slot_view(pair + 1, p->last_set);
    if (front_kps_un) FBE_CUDA(cudaMemcpy(front_kps_un, f.out_kps_un, (size_t)p->fcap * sizeof(fbe_keypoint), cudaMemcpyDeviceToHost));
    if (bounds) for (int i = 0; i < 4; ++i) bounds[i] = p->bounds[i];
    return FBE_OK;
}

int fbe_pipeline_fetch_pair(fbe_pipeline* p, int32_t pair, fbe_keypoint* front_kps, uint8_t* front_desc, fbe_keypoint* bird_kps,
                            uint8_t* bird_desc) {
    if (!p || pair < 0 || pair >= p->B) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(p->cfg.device));
    FBE_TRY(fbe_pipeline_sync(p));
    Workspace f = p->front.slot_view(pair + 1, p->last_set), b = p->bird.slot_view(pair + 1, p->last_set);
    if (front_kps) FBE_CUDA(cudaMemcpy(front_kps, f.out_kps, (size_t)p->fcap * sizeof(fbe_keypoint), cudaMemcpyDeviceToHost));
    if (front_desc) FBE_CUDA(cudaMemcpy(front_desc, f.out_desc, (size_t)p->fcap * 32, cudaMemcpyDeviceToHost));
    if (bird_kps) FBE_CUDA(cudaMemcpy(bird_kps, b.out_kps, (size_t)p->bcap * sizeof(fbe_keypoint), cudaMemcpyDeviceToHost));
    if (bird_desc) FBE_CUDA(cudaMemcpy(bird_desc, b.out_desc, (size_t)p->bcap * 32, cudaMemcpyDeviceToHost));
    return FBE_OK;
}

// Asynchronous step through HOST buffers (ideally pinned).  submit() enqueues the H2D copy of the inputs on a copy stream,
// the step, and the D2H copy of the results into the caller's buffers, then returns a ticket; wait(ticket) blocks until
// that step's results are in the host buffers.  Up to three steps may be in flight: the inputs of step N+1 travel while step N
// computes, so the end-to-end rate is max(copy, compute) instead of their sum.  Steps execute in submit order.
int fbe_pipeline_submit_host_features(fbe_pipeline* p, const uint8_t* h_front, const uint8_t* h_bird, fbe_pair_result* res,
                                      int32_t* front_matches12, int32_t* bird_matches12, const fbe_pipeline_features* feat,
                                      int32_t* ticket) {
    if (!p || !h_front || !h_bird || !ticket) return FBE_E_INVALID;
    const fbe_pipeline_cfg& c = p->cfg;
    FBE_CUDA(cudaSetDevice(c.device));
    const size_t B = (size_t)p->B;
    const size_t fbytes = B * c.front_rows * c.front_cols, bbytes = B * c.bird_rows * c.bird_cols;
    if (!p->copy_stream) {
        FBE_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
        FBE_CUDA(cudaMallocHost((void**)&p->h_flags, kFlagInts * fbe_pipeline::kQ * sizeof(int)));
        for (int k = 0; k < fbe_pipeline::kQ; ++k) {
            FBE_CUDA(cudaMalloc(&p->d_front_q[k], fbytes));
            FBE_CUDA(cudaMalloc(&p->d_bird_q[k], bbytes));
            FBE_CUDA(cudaEventCreate(&p->ev_in_ready[k]));
            FBE_CUDA(cudaEventCreate(&p->ev_copy0[k]));
            FBE_CUDA(cudaEventCreateWithFlags(&p->ev_in_free[k], cudaEventDisableTiming));
            FBE_CUDA(cudaEventCreateWithFlags(&p->ev_done[k], cudaEventDisableTiming));
        }
    }
    const int32_t t = p->next_ticket++;
    const int k = t % fbe_pipeline::kQ;
    cudaStream_t cs = p->copy_stream, ms = p->mstream;
    // inputs: wait until the step that last used this staging pair has consumed it
    if (p->in_used[k]) FBE_CUDA(cudaStreamWaitEvent(cs, p->ev_in_free[k], 0));
    FBE_CUDA(cudaEventRecord(p->ev_copy0[k], cs));
    FBE_CUDA(cudaMemcpyAsync(p->d_front_q[k], h_front, fbytes, cudaMemcpyHostToDevice, cs));
    FBE_CUDA(cudaMemcpyAsync(p->d_bird_q[k], h_bird, bbytes, cudaMemcpyHostToDevice, cs));
    FBE_CUDA(cudaEventRecord(p->ev_in_ready[k], cs));
    FBE_CUDA(cudaStreamWaitEvent(p->front.stream, p->ev_in_ready[k], 0));
    FBE_CUDA(cudaStreamWaitEvent(p->bird.stream, p->ev_in_ready[k], 0));
    FBE_TRY(fbe_pipeline_step_dev(p, p->d_front_q[k], p->d_bird_q[k]));
    const int set = p->last_set;
    if (feat && (feat->front_kps || feat->front_desc || feat->bird_kps || feat->bird_desc)) {
        // what ORBextractor::operator() hands back, for the batch frames (slots 1 .. B of this step's output set): copied on
        // its own stream as soon as each extractor is done, beside the matching; the set is not reused before ev_feat_done
        if (!p->out_stream) {
            FBE_CUDA(cudaStreamCreateWithFlags(&p->out_stream, cudaStreamNonBlocking));
            for (cudaEvent_t* e : {&p->ev_feat_done[0], &p->ev_feat_done[1], &p->ev_feat_q[0], &p->ev_feat_q[1], &p->ev_feat_q[2]})
                FBE_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        }
        cudaStream_t os = p->out_stream;
        const Workspace f = p->front.slot_view(1, set), b = p->bird.slot_view(1, set);
        FBE_CUDA(cudaStreamWaitEvent(os, p->ev_front, 0));
        if (feat->front_kps) FBE_CUDA(cudaMemcpyAsync(feat->front_kps, f.out_kps, B * p->fcap * sizeof(fbe_keypoint), cudaMemcpyDeviceToHost, os));
        if (feat->front_desc) FBE_CUDA(cudaMemcpyAsync(feat->front_desc, f.out_desc, B * p->fcap * 32, cudaMemcpyDeviceToHost, os));
        FBE_CUDA(cudaStreamWaitEvent(os, p->ev_bird, 0));
        if (feat->bird_kps) FBE_CUDA(cudaMemcpyAsync(feat->bird_kps, b.out_kps, B * p->bcap * sizeof(fbe_keypoint), cudaMemcpyDeviceToHost, os));
        if (feat->bird_desc) FBE_CUDA(cudaMemcpyAsync(feat->bird_desc, b.out_desc, B * p->bcap * 32, cudaMemcpyDeviceToHost, os));
        FBE_CUDA(cudaEventRecord(p->ev_feat_done[set], os));
        FBE_CUDA(cudaEventRecord(p->ev_feat_q[k], os));
        p->feat_used[set] = true;
        p->feat_q_used[k] = true;
    } else {
        p->feat_q_used[k] = false;
    }
    FBE_CUDA(cudaEventRecord(p->ev_in_free[k], ms));       // the match stream has joined both extractor streams inside step_dev
    p->in_used[k] = true;
    // results
    if (res) FBE_CUDA(cudaMemcpyAsync(res, p->d_res, B * sizeof(fbe_pair_result), cudaMemcpyDeviceToHost, ms));
    FBE_CUDA(cudaMemcpyAsync(p->h_flags + kFlagInts * k, p->flags, kFlagInts * sizeof(int), cudaMemcpyDeviceToHost, ms));
    if (front_matches12) FBE_CUDA(cudaMemcpyAsync(front_matches12, p->f_m12, B * p->fcap * 4, cudaMemcpyDeviceToHost, ms));
    if (bird_matches12) FBE_CUDA(cudaMemcpyAsync(bird_matches12, p->b_m12, B * p->bcap * 4, cudaMemcpyDeviceToHost, ms));
    FBE_CUDA(cudaEventRecord(p->ev_done[k], ms));
    // the next step's matching overwrites f_m12 / b_m12 / d_res: it is enqueued on ms after these copies, so ordering holds
    *ticket = t;
    return FBE_OK;
}

int fbe_pipeline_submit_host(fbe_pipeline* p, const uint8_t* h_front, const uint8_t* h_bird, fbe_pair_result* res,
                             int32_t* front_matches12, int32_t* bird_matches12, int32_t* ticket) {
    return fbe_pipeline_submit_host_features(p, h_front, h_bird, res, front_matches12, bird_matches12, nullptr, ticket);
}

int fbe_pipeline_device_results(fbe_pipeline* p, void** res, void** front_matches12, void** bird_matches12) {
    if (!p) return FBE_E_INVALID;
    if (res) *res = p->d_res;
    if (front_matches12) *front_matches12 = p->f_m12;
    if (bird_matches12) *bird_matches12 = p->b_m12;
    return FBE_OK;
}

// device milliseconds the input copy of a completed ticket took (diagnostic for the end-to-end number)
int fbe_pipeline_copy_ms(fbe_pipeline* p, int32_t ticket, float* ms) {
    if (!p || !ms || ticket < 0 || ticket >= p->next_ticket || ticket < p->next_ticket - fbe_pipeline::kQ) return FBE_E_INVALID;
    const int k = ticket % fbe_pipeline::kQ;
    FBE_CUDA(cudaEventSynchronize(p->ev_in_ready[k]));
    FBE_CUDA(cudaEventElapsedTime(ms, p->ev_copy0[k], p->ev_in_ready[k]));
    return FBE_OK;
}

int fbe_pipeline_wait(fbe_pipeline* p, int32_t ticket) {
    if (!p || ticket < 0 || ticket >= p->next_ticket || ticket < p->next_ticket - fbe_pipeline::kQ) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(p->cfg.device));
    const int k = ticket % fbe_pipeline::kQ;
    FBE_CUDA(cudaEventSynchronize(p->ev_done[k]));
    if (p->feat_q_used[k]) FBE_CUDA(cudaEventSynchronize(p->ev_feat_q[k]));
    return check_flags(p->h_flags + kFlagInts * k, p->row_cap);
}

// live per-stage device timing of the FRONT and BIRD extractors (CUDA events on their launching streams).
// ms: 12 doubles = front[pyramid, fast, octree, describe, grid, blur], bird[...]; steps: number of steps summed.
int fbe_pipeline_stage_timing(fbe_pipeline* p, int32_t enable) {
    if (!p) return FBE_E_INVALID;
    FBE_TRY(p->front.enable_timing(enable != 0));
    return p->bird.enable_timing(enable != 0);
}
int fbe_pipeline_stage_ms(fbe_pipeline* p, double* ms, int32_t* steps) {
    if (!p || !ms || !steps) return FBE_E_INVALID;
    FBE_TRY(fbe_pipeline_sync(p));
    int n1 = 0, n2 = 0;
    FBE_TRY(p->front.collect_timing(ms, &n1));
    FBE_TRY(p->bird.collect_timing(ms + 6, &n2));
    *steps = n1;
    return FBE_OK;
}

// pinned host memory for callers that have no CUDA runtime of their own (bench.py's host-buffer leg)
int fbe_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return FBE_E_INVALID;
    FBE_CUDA(cudaMallocHost(ptr, bytes));
    return FBE_OK;
}
int fbe_host_free(void* ptr) {
    if (ptr) FBE_CUDA(cudaFreeHost(ptr));
    return FBE_OK;
}

}  // extern "C"
