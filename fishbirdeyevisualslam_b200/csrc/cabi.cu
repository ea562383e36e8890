// C-ABI entry points for the extractor and the grid (include/fbe_cabi.h).  Host buffers in, host buffers out;
// everything in between is device work on the handle's stream.
#include <cstring>
#include <new>
#include "fbe_internal.cuh"

using namespace fbe;

#define FBE_TRY(expr) do { int _rc = (expr); if (_rc != FBE_OK) return _rc; } while (0)

struct fbe_extractor {
    ExtractorCore core;
};

extern "C" {

const char* fbe_last_error(void) { return last_error(); }
int fbe_version(void) { return 100; }
uint64_t fbe_kernel_launch_count(void) { return (uint64_t)g_launches.load(); }

int fbe_extractor_create(const fbe_extractor_cfg* cfg, fbe_extractor** out) {
    if (!cfg || !out) return FBE_E_INVALID;
    *out = nullptr;
    fbe_extractor* e = new (std::nothrow) fbe_extractor();
    if (!e) return FBE_E_INVALID;
    int rc = e->core.init(*cfg);
    if (rc != FBE_OK) { e->core.destroy(); delete e; return rc; }
    *out = e;
    return FBE_OK;
}

int fbe_extractor_destroy(fbe_extractor* e) {
    if (!e) return FBE_E_INVALID;
    e->core.destroy();
    delete e;
    return FBE_OK;
}

int fbe_extractor_tables(const fbe_extractor* e, int32_t* nlevels, const float** scale, const float** inv_scale,
                         const float** sigma2, const float** inv_sigma2) {
    if (!e) return FBE_E_INVALID;
    if (nlevels) *nlevels = e->core.cfg.nlevels;
    if (scale) *scale = e->core.scale.data();
    if (inv_scale) *inv_scale = e->core.inv_scale.data();
    if (sigma2) *sigma2 = e->core.sigma2.data();
    if (inv_sigma2) *inv_sigma2 = e->core.inv_sigma2.data();
    return FBE_OK;
}

int fbe_extractor_features_per_level(const fbe_extractor* e, const int32_t** per_level) {
    if (!e || !per_level) return FBE_E_INVALID;
    *per_level = e->core.per_level.data();
    return FBE_OK;
}

int fbe_extractor_max_keypoints(const fbe_extractor* e, int32_t rows, int32_t cols, int32_t* cap) {
    if (!e || !cap) return FBE_E_INVALID;
    Plan p;
    std::vector<ResizeTab> tabs;
    int rc = build_plan(e->core.cfg, e->core.scale, e->core.inv_scale, e->core.per_level, e->core.umax, rows, cols, p, tabs);
    if (rc != FBE_OK) return rc;
    *cap = p.kp_cap_total;
    return FBE_OK;
}

int fbe_plan_query(const fbe_extractor_cfg* cfg, int32_t rows, int32_t cols, int32_t* out, float* scale, float* inv_scale,
                   float* sigma2, float* inv_sigma2) {
    if (!cfg || !out || cfg->nlevels < 1 || cfg->nlevels > FBE_MAX_LEVELS || cfg->nfeatures < 1 || !(cfg->scale_factor > 1.0f)) return FBE_E_INVALID;
    std::vector<float> sc, isc, s2, is2;
    std::vector<int> per;
    int umax[16];
    compute_extractor_tables(cfg->nfeatures, cfg->scale_factor, cfg->nlevels, sc, isc, s2, is2, per, umax);
    Plan p;
    std::vector<ResizeTab> tabs;
    int rc = build_plan(*cfg, sc, isc, per, umax, rows, cols, p, tabs);
    if (rc != FBE_OK) return rc;
    for (int l = 0; l < cfg->nlevels; ++l) {
        const LevelGeom& g = p.lv[l];
        int32_t* o = out + 8 * l;
        o[0] = g.w; o[1] = g.h; o[2] = g.ncols; o[3] = g.nrows; o[4] = g.wcell; o[5] = g.hcell; o[6] = g.nfeat; o[7] = g.nini;
        if (scale) scale[l] = sc[l];
        if (inv_scale) inv_scale[l] = isc[l];
        if (sigma2) sigma2[l] = s2[l];
        if (inv_sigma2) inv_sigma2[l] = is2[l];
    }
    return FBE_OK;
}

// pyr_dst / pyr_step (optional, one image only): the padded pyramid levels are copied out on the side stream as soon as the pyramid
// kernels have finished, i.e. behind FAST / octree / describe instead of after them.
static int extract_batch_impl(fbe_extractor* e, const uint8_t* const* imgs, int32_t nimg, int32_t rows, int32_t cols, size_t step,
                              fbe_keypoint* kps, uint8_t* desc, int32_t capacity, int32_t* n_out, uint8_t* const* pyr_dst,
                              const size_t* pyr_step) {
    if (!e || !n_out) return FBE_E_INVALID;
    for (int i = 0; i < nimg; ++i) n_out[i] = 0;
    if (!imgs || nimg <= 0 || rows <= 0 || cols <= 0) return FBE_OK;   // reference: empty image -> silent return (:1046-1047)
    for (int i = 0; i < nimg; ++i) if (!imgs[i]) return FBE_OK;
    if (step < (size_t)cols || capacity < 0 || (capacity > 0 && (!kps || !desc))) return FBE_E_INVALID;
    ExtractorCore& c = e->core;
    FBE_CUDA(cudaSetDevice(c.cfg.device));
    int rc = c.ensure_plan(rows, cols);
    if (rc != FBE_OK) return rc;
    const int cap_total = c.hplan.kp_cap_total;
    const size_t img_bytes = (size_t)rows * cols;
    for (int base = 0; base < nimg; base += c.cfg.max_batch) {
        const int nb = std::min(c.cfg.max_batch, nimg - base);
        if (c.d_in_bytes < img_bytes * c.cfg.max_batch) {
            cudaFree(c.d_in); c.d_in = nullptr;
            FBE_CUDA(cudaMalloc(&c.d_in, img_bytes * c.cfg.max_batch));
            c.d_in_bytes = img_bytes * c.cfg.max_batch;
        }
        const size_t out_bytes = (size_t)c.cfg.max_batch * ((size_t)cap_total * (sizeof(fbe_keypoint) + 32) + 2 * sizeof(int));
        if (c.h_pin_bytes < out_bytes) {
            if (c.h_pin) cudaFreeHost(c.h_pin);
            c.h_pin = nullptr;
            FBE_CUDA(cudaMallocHost(&c.h_pin, out_bytes));
            c.h_pin_bytes = out_bytes;
        }
        for (int i = 0; i < nb; ++i)
            FBE_CUDA(cudaMemcpy2DAsync(c.d_in + (size_t)i * img_bytes, cols, imgs[base + i], step, cols, rows,
                                       cudaMemcpyHostToDevice, c.stream));
        rc = c.run_dev(c.d_in, cols, (int)img_bytes, nb, rows, cols);
        if (rc != FBE_OK) return rc;
        if (pyr_dst) {                 // stream2 has run the blur behind the pyramid event; the level copies queue up after it
            for (int l = 0; l < c.cfg.nlevels; ++l) {
                const LevelGeom& g = c.hplan.lv[l];
                FBE_CUDA(cudaMemcpy2DAsync(pyr_dst[l], pyr_step[l], c.ws.pyr + g.img_off, g.pitch, g.w + 2 * kEdge, g.ph,
                                           cudaMemcpyDeviceToHost, c.stream2));
            }
        }
        uint8_t* hp = c.h_pin;
        fbe_keypoint* h_kps = reinterpret_cast<fbe_keypoint*>(hp);
        uint8_t* h_desc = hp + (size_t)c.cfg.max_batch * cap_total * sizeof(fbe_keypoint);
        int* h_n = reinterpret_cast<int*>(h_desc + (size_t)c.cfg.max_batch * cap_total * 32);
        int* h_status = h_n + c.cfg.max_batch;
        FBE_CUDA(cudaMemcpyAsync(h_n, c.ws.out_n, nb * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        FBE_CUDA(cudaMemcpyAsync(h_status, c.ws.status, nb * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        FBE_CUDA(cudaMemcpyAsync(h_kps, c.ws.out_kps, (size_t)nb * cap_total * sizeof(fbe_keypoint), cudaMemcpyDeviceToHost, c.stream));
        FBE_CUDA(cudaMemcpyAsync(h_desc, c.ws.out_desc, (size_t)nb * cap_total * 32, cudaMemcpyDeviceToHost, c.stream));
        FBE_CUDA(cudaStreamSynchronize(c.stream));
        if (pyr_dst) FBE_CUDA(cudaStreamSynchronize(c.stream2));
        for (int i = 0; i < nb; ++i) {
            if (h_status[i]) { set_error("octree workspace overflow"); return FBE_E_CAPACITY; }
            const int n = h_n[i];
            n_out[base + i] = n;
            if (n > capacity) { set_error("keypoint capacity too small"); return FBE_E_CAPACITY; }
            if (n > 0) {
                std::memcpy(kps + (size_t)(base + i) * capacity, h_kps + (size_t)i * cap_total, (size_t)n * sizeof(fbe_keypoint));
                std::memcpy(desc + (size_t)(base + i) * capacity * 32, h_desc + (size_t)i * cap_total * 32, (size_t)n * 32);
            }
        }
    }
    return FBE_OK;
}

int fbe_extract_batch(fbe_extractor* e, const uint8_t* const* imgs, int32_t nimg, int32_t rows, int32_t cols, size_t step,
                      fbe_keypoint* kps, uint8_t* desc, int32_t capacity, int32_t* n_out) {
    return extract_batch_impl(e, imgs, nimg, rows, cols, step, kps, desc, capacity, n_out, nullptr, nullptr);
}

int fbe_extract_pyramid(fbe_extractor* e, const uint8_t* img, int32_t rows, int32_t cols, size_t step, fbe_keypoint* kps, uint8_t* desc,
                        int32_t capacity, int32_t* n_out, uint8_t* const* pyr_dst, const size_t* pyr_step) {
    if (!e || !pyr_dst || !pyr_step) return FBE_E_INVALID;
    if (!img || rows <= 0 || cols <= 0) { if (n_out) *n_out = 0; return n_out ? FBE_OK : FBE_E_INVALID; }
    FBE_CUDA(cudaSetDevice(e->core.cfg.device));
    int rc = e->core.ensure_plan(rows, cols);          // the level sizes the destinations are checked against
    if (rc != FBE_OK) return rc;
    for (int l = 0; l < e->core.cfg.nlevels; ++l)
        if (!pyr_dst[l] || pyr_step[l] < (size_t)(e->core.hplan.lv[l].w + 2 * kEdge)) return FBE_E_INVALID;
    const uint8_t* one[1] = {img};
    return extract_batch_impl(e, one, 1, rows, cols, step, kps, desc, capacity, n_out, pyr_dst, pyr_step);
}

int fbe_extract(fbe_extractor* e, const uint8_t* img, int32_t rows, int32_t cols, size_t step, fbe_keypoint* kps,
                uint8_t* desc, int32_t capacity, int32_t* n_out) {
    const uint8_t* one[1] = {img};
    return fbe_extract_batch(e, one, 1, rows, cols, step, kps, desc, capacity, n_out);
}

int fbe_pyramid_level(fbe_extractor* e, int32_t slot, int32_t level, uint8_t* dst, size_t dst_step, int32_t* rows, int32_t* cols) {
    if (!e || !e->core.have_ws || level < 0 || level >= e->core.cfg.nlevels || slot < 0 || slot >= e->core.cfg.max_batch) return FBE_E_INVALID;
    ExtractorCore& c = e->core;
    const LevelGeom& g = c.hplan.lv[level];
    if (rows) *rows = g.h;
    if (cols) *cols = g.w;
    if (!dst) return FBE_OK;
    FBE_CUDA(cudaSetDevice(c.cfg.device));
    FBE_CUDA(cudaMemcpy2DAsync(dst, dst_step, c.ws.pyr + (size_t)slot * c.hplan.pyr_bytes + g.img_off, g.pitch, g.w + 2 * kEdge, g.ph,
                               cudaMemcpyDeviceToHost, c.stream));
    FBE_CUDA(cudaStreamSynchronize(c.stream));
    return FBE_OK;
}

int fbe_pyramid_geometry(fbe_extractor* e, int32_t rows, int32_t cols, int32_t* level_rows, int32_t* level_cols) {
    if (!e || !level_rows || !level_cols || rows <= 0 || cols <= 0) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(e->core.cfg.device));
    const int rc = e->core.ensure_plan(rows, cols);
    if (rc != FBE_OK) return rc;
    for (int l = 0; l < e->core.cfg.nlevels; ++l) { level_rows[l] = e->core.hplan.lv[l].h; level_cols[l] = e->core.hplan.lv[l].w; }
    return FBE_OK;
}

int fbe_pyramid_fetch(fbe_extractor* e, int32_t slot, uint8_t* const* dst, const size_t* dst_step) {
    if (!e || !dst || !dst_step || !e->core.have_ws || slot < 0 || slot >= e->core.cfg.max_batch) return FBE_E_INVALID;
    ExtractorCore& c = e->core;
    FBE_CUDA(cudaSetDevice(c.cfg.device));
    for (int l = 0; l < c.cfg.nlevels; ++l) {
        const LevelGeom& g = c.hplan.lv[l];
        if (!dst[l] || dst_step[l] < (size_t)(g.w + 2 * kEdge)) return FBE_E_INVALID;
        FBE_CUDA(cudaMemcpy2DAsync(dst[l], dst_step[l], c.ws.pyr + (size_t)slot * c.hplan.pyr_bytes + g.img_off, g.pitch, g.w + 2 * kEdge, g.ph,
                                   cudaMemcpyDeviceToHost, c.stream));
    }
    FBE_CUDA(cudaStreamSynchronize(c.stream));
    return FBE_OK;
}

int fbe_debug_blurred(fbe_extractor* e, int32_t slot, int32_t level, uint8_t* dst, int32_t* rows, int32_t* cols) {
    if (!e || !e->core.have_ws || level < 0 || level >= e->core.cfg.nlevels || slot < 0 || slot >= e->core.cfg.max_batch) return FBE_E_INVALID;
    ExtractorCore& c = e->core;
    const LevelGeom& g = c.hplan.lv[level];
    if (rows) *rows = g.h;
    if (cols) *cols = g.w;
    if (!dst) return FBE_OK;
    FBE_CUDA(cudaSetDevice(c.cfg.device));
    FBE_CUDA(cudaMemcpy2DAsync(dst, g.w, c.ws.blur + (size_t)slot * c.hplan.pyr_bytes + g.img_off + (size_t)kEdge * g.pitch + kEdge,
                               g.pitch, g.w, g.h, cudaMemcpyDeviceToHost, c.stream));
    FBE_CUDA(cudaStreamSynchronize(c.stream));
    return FBE_OK;
}

int fbe_debug_candidates(fbe_extractor* e, int32_t slot, int32_t level, int32_t* xys, int32_t cap, int32_t* n) {
    if (!e || !n || !e->core.have_ws || level < 0 || level >= e->core.cfg.nlevels || slot < 0 || slot >= e->core.cfg.max_batch) return FBE_E_INVALID;
    ExtractorCore& c = e->core;
    const LevelGeom& g = c.hplan.lv[level];
    FBE_CUDA(cudaSetDevice(c.cfg.device));
    const int ncells = g.ncols * g.nrows;
    std::vector<int> cnt(ncells);
    std::vector<uint32_t> slots((size_t)g.key_cap);
    FBE_CUDA(cudaMemcpyAsync(cnt.data(), c.ws.cell_count + (size_t)slot * c.hplan.ncells_total + g.cell_base, ncells * sizeof(int),
                             cudaMemcpyDeviceToHost, c.stream));
    FBE_CUDA(cudaMemcpyAsync(slots.data(), c.ws.slots + (size_t)slot * c.hplan.slots_total + g.slot_base, slots.size() * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, c.stream));
    FBE_CUDA(cudaStreamSynchronize(c.stream));
    int total = 0;
    for (int ci = 0; ci < ncells; ++ci) {
        for (int i = 0; i < cnt[ci]; ++i, ++total) {
            if (xys && total < cap) {
                const uint32_t k = slots[(size_t)ci * g.cell_cap + i];
                xys[3 * total] = key_x(k); xys[3 * total + 1] = key_y(k); xys[3 * total + 2] = key_s(k);
            }
        }
    }
    *n = total;
    return FBE_OK;
}

int fbe_debug_octree(const int32_t* xys, int32_t n, int32_t min_x, int32_t max_x, int32_t min_y, int32_t max_y, int32_t nfeat,
                     int32_t* sel, int32_t cap, int32_t* n_sel) {
    if (!n_sel || n < 0 || (n > 0 && !xys) || max_x <= min_x || max_y <= min_y) return FBE_E_INVALID;
    const int W = max_x - min_x, H = max_y - min_y;
    const int nini = (int)roundf((float)W / (float)H);
    if (nini <= 0) { set_error("zero octree roots"); return FBE_E_UNSUPPORTED; }
    const float hx = (float)W / (float)nini;
    const int ncap = std::max(nfeat, 4 * nini) + 8;
    std::vector<uint32_t> keys(std::max(n, 1));
    for (int i = 0; i < n; ++i) {
        if (xys[3 * i] < 0 || xys[3 * i] + 16 > kMaxDim || xys[3 * i + 1] < 0 || xys[3 * i + 1] + 16 > kMaxDim) return FBE_E_INVALID;
        keys[i] = pack_key(xys[3 * i] + 16, xys[3 * i + 1] + 16, xys[3 * i + 2]);
    }
    // per-thread arena (nothing to leak on an error return, no cudaMalloc / cudaFree in the steady state)
    static thread_local Arena arena;
    int dev = 0;
    FBE_CUDA(cudaGetDevice(&dev));
    const size_t scr = octree_debug_scratch_bytes(ncap, nini, nfeat);
    FBE_CUDA(arena.reserve(2 * pad256(keys.size() * 4) + pad256((size_t)ncap * 4) + pad256(scr) + pad256(4), dev));
    uint32_t* d_keys = arena.take<uint32_t>(keys.size());
    uint32_t* d_knode = arena.take<uint32_t>(keys.size());
    uint32_t* d_sel = arena.take<uint32_t>((size_t)ncap);
    uint8_t* d_scr = arena.take<uint8_t>(scr);
    int* d_n = arena.take<int>(1);
    FBE_CUDA(cudaMemcpy(d_keys, keys.data(), keys.size() * 4, cudaMemcpyHostToDevice));
    int rc = launch_octree_debug(d_keys, d_knode, n, nini, hx, H, nfeat, ncap, d_scr, d_sel, d_n, 0);
    int hn = 0;
    std::vector<uint32_t> hsel(ncap);
    if (rc == FBE_OK) {
        cudaError_t ce = cudaMemcpy(&hn, d_n, 4, cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess && hn > 0) ce = cudaMemcpy(hsel.data(), d_sel, (size_t)hn * 4, cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) { set_error(cudaGetErrorString(ce)); rc = FBE_E_CUDA; }
    }
    if (rc != FBE_OK) return rc;
    if (hn < 0) { set_error("octree workspace overflow"); return FBE_E_CAPACITY; }
    *n_sel = hn;
    for (int i = 0; i < hn && i < cap; ++i) sel[i] = (int32_t)hsel[i];
    return FBE_OK;
}

int fbe_set_device(int32_t device) {
    FBE_CUDA(cudaSetDevice(device));
    return FBE_OK;
}

int fbe_grid_assign(const fbe_keypoint* kps, int32_t n, float min_x, float min_y, float inv_w, float inv_h, int32_t gcols,
                    int32_t grows, int32_t* cell_start, int32_t* cell_items, int32_t* n_assigned) {
    if (n < 0 || gcols <= 0 || grows <= 0 || !cell_start || (n > 0 && (!kps || !cell_items))) return FBE_E_INVALID;
    const int gcells = gcols * grows;
    const size_t nn = (size_t)std::max(n, 1);
    // per-thread arena on the thread's current device (fbe_set_device): no cudaMalloc / cudaFree in the steady state, nothing
    // to leak on an error return
    static thread_local Arena arena;
    int dev = 0;
    FBE_CUDA(cudaGetDevice(&dev));
    FBE_CUDA(arena.reserve(pad256(nn * sizeof(fbe_keypoint)) + pad256(4) + 2 * pad256(nn * 4) + pad256((size_t)(gcells + 1) * 4), dev));
    fbe_keypoint* d_kps = arena.take<fbe_keypoint>(nn);
    int* d_n = arena.take<int>(1);
    int* d_cell = arena.take<int>(nn);
    int* d_start = arena.take<int>((size_t)gcells + 1);
    int* d_items = arena.take<int>(nn);
    cudaStream_t st = cudaStreamPerThread;
    if (n > 0) FBE_CUDA(cudaMemcpyAsync(d_kps, kps, (size_t)n * sizeof(fbe_keypoint), cudaMemcpyHostToDevice, st));
    FBE_CUDA(cudaMemcpyAsync(d_n, &n, 4, cudaMemcpyHostToDevice, st));
    FBE_TRY(launch_grid_build(d_kps, d_n, (int)nn, 1, min_x, min_y, inv_w, inv_h, gcols, grows, d_cell, d_start, d_items, st));
    FBE_CUDA(cudaMemcpyAsync(cell_start, d_start, (size_t)(gcells + 1) * 4, cudaMemcpyDeviceToHost, st));
    // every keypoint lands in at most one cell: the whole item array is at most n entries (one copy, one synchronisation)
    if (n > 0) FBE_CUDA(cudaMemcpyAsync(cell_items, d_items, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    FBE_CUDA(cudaStreamSynchronize(st));
    if (n_assigned) *n_assigned = cell_start[gcells];
    return FBE_OK;
}

int fbe_hamming256(const uint8_t a[32], const uint8_t b[32]) {
    int d = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t x, y;
        std::memcpy(&x, a + 8 * i, 8);
        std::memcpy(&y, b + 8 * i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

}  // extern "C"
