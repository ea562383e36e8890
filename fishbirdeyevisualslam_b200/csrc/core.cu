// Extractor core: owns the plan, the device workspace and the stream; enqueues the kernel sequence
//   level0/resize x (L-1)  ->  FAST cells  ->  octree  ->  [blur on a 2nd stream]  ->  orient+describe  ->  grid
// for a batch of equally sized images.  No host round trip inside the sequence.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "fbe_internal.cuh"
#include "tma.cuh"

namespace fbe {

static thread_local std::string t_err;
void set_error(const std::string& s) { t_err = s; }
const char* last_error() { return t_err.c_str(); }
std::atomic<unsigned long long> g_launches{0};

int ExtractorCore::init(const fbe_extractor_cfg& c) {
    cfg = c;
    if (cfg.nlevels < 1 || cfg.nlevels > FBE_MAX_LEVELS || cfg.nfeatures < 1 || !(cfg.scale_factor > 1.0f) ||
        cfg.ini_th_fast < 0 || cfg.min_th_fast < 0 || cfg.ini_th_fast > 254 || cfg.min_th_fast > 254) {
        set_error("invalid extractor configuration");
        return FBE_E_INVALID;
    }
    if (cfg.max_batch < 1) cfg.max_batch = 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("no CUDA device: this library has no CPU path");
        return FBE_E_CUDA;
    }
    FBE_CUDA(cudaSetDevice(cfg.device));
    compute_extractor_tables(cfg.nfeatures, cfg.scale_factor, cfg.nlevels, scale, inv_scale, sigma2, inv_sigma2, per_level, umax);
    FBE_CUDA(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, stream_priority));
    FBE_CUDA(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
    FBE_CUDA(cudaEventCreateWithFlags(&ev_pyr, cudaEventDisableTiming));
    FBE_CUDA(cudaEventCreateWithFlags(&ev_blur, cudaEventDisableTiming));
    std::memset(&ws, 0, sizeof(ws));
    return FBE_OK;
}

int ExtractorCore::free_ws() {
    cudaFree(ws.pyr); cudaFree(ws.blur); cudaFree(ws.cell_count); cudaFree(ws.slots); cudaFree(ws.keys);
    cudaFree(ws.key_node); cudaFree(ws.oct_scratch); cudaFree(ws.sel); cudaFree(ws.level_n); if (ws.out_kps_un != ws.out_kps) cudaFree(ws.out_kps_un);
    cudaFree(ws.out_kps);
    cudaFree(ws.out_desc); cudaFree(ws.out_n); cudaFree(ws.out_cell); cudaFree(ws.grid_start); cudaFree(ws.grid_items);
    cudaFree(ws.status); cudaFree(dplan); cudaFree(dtab);
    cudaFree(const_cast<uint32_t*>(ws.fast_tab)); cudaFree(const_cast<uint32_t*>(ws.blur_tab));
    delete blur_maps; blur_maps = nullptr;
    delete fast_maps; fast_maps = nullptr;
    delete rs_maps; rs_maps = nullptr;
    std::memset(&ws, 0, sizeof(ws));
    dplan = nullptr; dtab = nullptr; have_ws = false;
    return FBE_OK;
}

void ExtractorCore::destroy() {
    cudaSetDevice(cfg.device);
    if (stream) cudaStreamSynchronize(stream);
    if (stream2) cudaStreamSynchronize(stream2);
    free_ws();
    cudaFree(d_in); d_in = nullptr;
    if (h_pin) cudaFreeHost(h_pin);
    h_pin = nullptr;
    for (auto& e : tev) cudaEventDestroy(e);
    tev.clear();
    if (ev_pyr) cudaEventDestroy(ev_pyr);
    if (ev_blur) cudaEventDestroy(ev_blur);
    if (stream) cudaStreamDestroy(stream);
    if (stream2) cudaStreamDestroy(stream2);
    stream = stream2 = nullptr;
}

int ExtractorCore::set_fisheye(const float K[4], const float D[4]) {
    if (!K || !D) return FBE_E_INVALID;
    FBE_CUDA(cudaSetDevice(cfg.device));
    if (have_ws) { FBE_CUDA(cudaStreamSynchronize(stream)); FBE_CUDA(cudaStreamSynchronize(stream2)); free_ws(); have_ws = false; }
    fisheye = true;
    for (int i = 0; i < 4; ++i) { fish_K[i] = K[i]; fish_D[i] = D[i]; }
    return FBE_OK;
}

int ExtractorCore::set_grid(float min_x, float min_y, float inv_w, float inv_h, int gcols, int grows) {
    // the same limit as the grid kernel's shared-memory histogram (launch_grid_build): a grid accepted here always runs
    if (gcols < 0 || grows < 0 || (size_t)(gcols * grows + 1) * sizeof(int) > 48 * 1024) { set_error("grid too large"); return FBE_E_UNSUPPORTED; }
    const bool realloc_grid = have_ws && gcols * grows != g_cols * g_rows;
    g_min_x = min_x; g_min_y = min_y; g_inv_w = inv_w; g_inv_h = inv_h; g_cols = gcols; g_rows = grows;
    if (have_ws) {
        hplan.grid_min_x = min_x; hplan.grid_min_y = min_y; hplan.grid_inv_w = inv_w; hplan.grid_inv_h = inv_h;
        hplan.grid_cols = gcols; hplan.grid_rows = grows;
        FBE_CUDA(cudaMemcpyAsync(dplan, &hplan, sizeof(Plan), cudaMemcpyHostToDevice, stream));
        FBE_CUDA(cudaStreamSynchronize(stream));
        if (realloc_grid) {
            FBE_CUDA(cudaDeviceSynchronize());            // every stream that may still read the old CSR (matching, copy-out) has drained
            cudaFree(ws.grid_start);
            ws.grid_start = nullptr;                      // never left dangling if the allocation below fails
            FBE_CUDA(cudaMalloc(&ws.grid_start, (size_t)cfg.max_batch * out_sets * (std::max(gcols * grows, kMaxGridCells) + 1) * sizeof(int)));
        }
    }
    return FBE_OK;
}

int ExtractorCore::ensure_plan(int rows, int cols) {
    if (have_ws && rows == plan_rows && cols == plan_cols) return FBE_OK;
    const int rc = build_workspace(rows, cols);
    if (rc != FBE_OK) free_ws();                          // a failed allocation leaves nothing behind (free_ws accepts null members)
    return rc;
}

int ExtractorCore::build_workspace(int rows, int cols) {
    FBE_CUDA(cudaSetDevice(cfg.device));
    if (have_ws) { FBE_CUDA(cudaStreamSynchronize(stream)); FBE_CUDA(cudaStreamSynchronize(stream2)); free_ws(); }
    std::vector<ResizeTab> tabs;
    int rc = build_plan(cfg, scale, inv_scale, per_level, umax, rows, cols, hplan, tabs);
    if (rc != FBE_OK) return rc;
    hplan.grid_min_x = g_min_x; hplan.grid_min_y = g_min_y; hplan.grid_inv_w = g_inv_w; hplan.grid_inv_h = g_inv_h;
    hplan.grid_cols = g_cols; hplan.grid_rows = g_rows;
    const size_t B = (size_t)cfg.max_batch;
    FBE_CUDA(cudaMalloc(&dplan, sizeof(Plan)));
    FBE_CUDA(cudaMemcpy(dplan, &hplan, sizeof(Plan), cudaMemcpyHostToDevice));
    FBE_CUDA(cudaMalloc(&dtab, std::max<size_t>(tabs.size(), 1) * sizeof(ResizeTab)));
    if (!tabs.empty()) FBE_CUDA(cudaMemcpy(dtab, tabs.data(), tabs.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice));
    FBE_CUDA(cudaMalloc(&ws.pyr, B * hplan.pyr_bytes));
    FBE_CUDA(cudaMalloc(&ws.blur, B * hplan.pyr_bytes));
    blur_maps = new TmaMaps();
    fast_maps = new TmaMaps();
    rs_maps = new TmaMaps();
    std::memset(rs_maps, 0, sizeof(TmaMaps));
    std::memset(blur_maps, 0, sizeof(TmaMaps));
    std::memset(fast_maps, 0, sizeof(TmaMaps));
    for (int l = 0; l < hplan.nlevels; ++l) {
        const LevelGeom& g = hplan.lv[l];
        rc = tma_encode_level(&blur_maps->m[l], ws.pyr + g.img_off, g.pitch, g.ph, (int)B, (size_t)hplan.pyr_bytes, kBlurTW + 32, kBlurTH + 6);
        if (rc != FBE_OK) return rc;
        rc = tma_encode_level(&fast_maps->m[l], ws.pyr + g.img_off, g.pitch, g.ph, (int)B, (size_t)hplan.pyr_bytes, 256, g.hcell + 6);
        if (rc != FBE_OK) return rc;
        if (l > 0 && g.rs_bw > 0) {
            const LevelGeom& sg = hplan.lv[l - 1];
            rc = tma_encode_level(&rs_maps->m[l], ws.pyr + sg.img_off, sg.pitch, sg.ph, (int)B, (size_t)hplan.pyr_bytes, g.rs_bw, g.rs_bh);
            if (rc != FBE_OK) return rc;
        }
    }
    FBE_CUDA(cudaMemset(ws.blur, 0, B * hplan.pyr_bytes));
    FBE_CUDA(cudaMalloc(&ws.cell_count, B * hplan.ncells_total * sizeof(int)));
    FBE_CUDA(cudaMalloc(&ws.slots, B * hplan.slots_total * sizeof(uint32_t)));
    FBE_CUDA(cudaMalloc(&ws.keys, B * hplan.slots_total * sizeof(uint32_t)));
    FBE_CUDA(cudaMalloc(&ws.key_node, B * hplan.slots_total * sizeof(uint32_t)));
    ws.oct_scratch_bytes = (octree_scratch_bytes(hplan) + 255) & ~(size_t)255;
    FBE_CUDA(cudaMalloc(&ws.oct_scratch, B * ws.oct_scratch_bytes));
    FBE_CUDA(cudaMalloc(&ws.sel, B * hplan.kp_cap_total * sizeof(uint32_t)));
    FBE_CUDA(cudaMalloc(&ws.level_n, B * FBE_MAX_LEVELS * sizeof(int)));
    FBE_CUDA(cudaMemset(ws.level_n, 0, B * FBE_MAX_LEVELS * sizeof(int)));
    const size_t Bo = B * (size_t)out_sets;           // output arrays exist once per output set
    FBE_CUDA(cudaMalloc(&ws.out_kps, Bo * hplan.kp_cap_total * sizeof(fbe_keypoint)));
    if (fisheye) FBE_CUDA(cudaMalloc(&ws.out_kps_un, Bo * hplan.kp_cap_total * sizeof(fbe_keypoint)));
    else ws.out_kps_un = ws.out_kps;
    FBE_CUDA(cudaMalloc(&ws.out_desc, Bo * hplan.kp_cap_total * 32));
    FBE_CUDA(cudaMalloc(&ws.out_n, Bo * sizeof(int)));
    FBE_CUDA(cudaMalloc(&ws.out_cell, Bo * hplan.kp_cap_total * sizeof(int)));
    const int gcells = std::max(g_cols * g_rows, kMaxGridCells);
    FBE_CUDA(cudaMalloc(&ws.grid_start, Bo * (gcells + 1) * sizeof(int)));
    FBE_CUDA(cudaMalloc(&ws.grid_items, Bo * hplan.kp_cap_total * sizeof(int)));
    FBE_CUDA(cudaMalloc(&ws.status, B * sizeof(int)));
    FBE_CUDA(cudaMemset(ws.status, 0, B * sizeof(int)));
    {
        std::vector<uint32_t> ft, bt;
        for (int l = 0; l < hplan.nlevels; ++l) {
            const LevelGeom& g = hplan.lv[l];
            for (int ci = 0; ci < g.nrows; ++ci)
                for (int gj = 0; gj < g.ngrp; ++gj) ft.push_back(((uint32_t)l << 24) | ((uint32_t)ci << 12) | (uint32_t)gj);
            for (int ty = 0; ty < g.blur_nty; ++ty)
                for (int tx = 0; tx < g.blur_ntx; ++tx) bt.push_back(((uint32_t)l << 24) | ((uint32_t)ty << 12) | (uint32_t)tx);
        }
        uint32_t *dft = nullptr, *dbt = nullptr;
        FBE_CUDA(cudaMalloc(&dft, ft.size() * 4));
        FBE_CUDA(cudaMalloc(&dbt, bt.size() * 4));
        FBE_CUDA(cudaMemcpy(dft, ft.data(), ft.size() * 4, cudaMemcpyHostToDevice));
        FBE_CUDA(cudaMemcpy(dbt, bt.data(), bt.size() * 4, cudaMemcpyHostToDevice));
        ws.fast_tab = dft; ws.blur_tab = dbt;
    }
    plan_rows = rows; plan_cols = cols; have_ws = true;
    return FBE_OK;
}

Workspace ExtractorCore::slot_view(int slot0, int out_set) const {
    Workspace v = ws;
    v.slot0 = slot0;
    const size_t s = (size_t)slot0;
    const size_t so = (size_t)out_set * cfg.max_batch + s;      // slot index inside the output arrays
    const int gcells = hplan.grid_cols * hplan.grid_rows;
    v.pyr += s * hplan.pyr_bytes; v.blur += s * hplan.pyr_bytes;
    v.cell_count += s * hplan.ncells_total;
    v.slots += s * hplan.slots_total; v.keys += s * hplan.slots_total; v.key_node += s * hplan.slots_total;
    v.oct_scratch += s * ws.oct_scratch_bytes;
    v.sel += s * hplan.kp_cap_total; v.level_n += s * FBE_MAX_LEVELS;
    v.out_kps += so * hplan.kp_cap_total; v.out_kps_un += so * hplan.kp_cap_total; v.out_desc += so * hplan.kp_cap_total * 32; v.out_n += so;
    v.out_cell += so * hplan.kp_cap_total; v.grid_start += so * (gcells + 1); v.grid_items += so * hplan.kp_cap_total;
    v.status += s;
    return v;
}

int ExtractorCore::run_dev(const uint8_t* d_imgs, int pitch, int slot_stride, int nimg, int rows, int cols, int slot0, int out_set) {
    if (nimg < 1 || slot0 < 0 || slot0 + nimg > cfg.max_batch || out_set < 0 || out_set >= out_sets) { set_error("batch larger than max_batch"); return FBE_E_INVALID; }
    FBE_CUDA(cudaSetDevice(cfg.device));
    int rc = ensure_plan(rows, cols);
    if (rc != FBE_OK) return rc;
    Workspace v = slot_view(slot0, out_set);
    v.in = d_imgs; v.in_pitch = pitch; v.in_slot_stride = slot_stride;
    cudaEvent_t* te = nullptr;
    if (timing) {
        te = tev.data() + (size_t)tpos * (kStages + 3);
        tpos = (tpos + 1) % kTimingRing;
        tcount = std::min(tcount + 1, kTimingRing);
    }
#define FBE_MARK(i, st) do { if (te) FBE_CUDA(cudaEventRecord(te[i], st)); } while (0)
    // FBE_SYNC_DEBUG=1: synchronise after every stage so that a device fault is attributed to the stage that raised it
    static const bool sync_debug = std::getenv("FBE_SYNC_DEBUG") != nullptr;
#define FBE_STAGE(name, st)                                                                              \
    do {                                                                                                 \
        if (sync_debug) {                                                                                \
            cudaError_t _e = cudaStreamSynchronize(st);                                                  \
            if (_e != cudaSuccess) { set_error(std::string("stage ") + name + ": " + cudaGetErrorString(_e)); return FBE_E_CUDA; } \
        }                                                                                                \
    } while (0)
    FBE_CUDA(cudaMemsetAsync(v.status, 0, (size_t)nimg * sizeof(int), stream));   // octree workspace-overflow flags of THIS run
    FBE_MARK(0, stream);
    if ((rc = launch_pyramid(hplan, dplan, v, dtab, *rs_maps, nimg, stream)) != FBE_OK) return rc;
    FBE_MARK(1, stream);
    FBE_STAGE("pyramid", stream);
    // blur depends only on the pyramid: run it on the side stream while FAST + octree proceed
    FBE_CUDA(cudaEventRecord(ev_pyr, stream));
    FBE_CUDA(cudaStreamWaitEvent(stream2, ev_pyr, 0));
    FBE_MARK(6, stream2);
    if ((rc = launch_blur(hplan, dplan, v, *blur_maps, nimg, stream2)) != FBE_OK) return rc;
    FBE_MARK(7, stream2);
    FBE_STAGE("blur", stream2);
    FBE_CUDA(cudaEventRecord(ev_blur, stream2));
    if ((rc = launch_fast_cells(hplan, dplan, v, *fast_maps, nimg, stream)) != FBE_OK) return rc;
    FBE_MARK(2, stream);
    FBE_STAGE("fast", stream);
    if ((rc = launch_octree(hplan, dplan, v, nimg, stream)) != FBE_OK) return rc;
    FBE_MARK(3, stream);
    FBE_STAGE("octree", stream);
    FBE_CUDA(cudaStreamWaitEvent(stream, ev_blur, 0));
    FBE_MARK(8, stream);
    if ((rc = launch_describe(hplan, dplan, v, nimg, stream)) != FBE_OK) return rc;
    FBE_MARK(4, stream);
    FBE_STAGE("describe", stream);
    if (fisheye && fish_D[0] != 0.0f) {            // Frame::UndistortKeyPoints (src/Frame.cc:638-669); k1 == 0 copies (:640-644)
        if ((rc = launch_undistort_batch(v.out_kps, v.out_n, hplan.kp_cap_total, nimg, fish_K, fish_D, v.out_kps_un, stream)) != FBE_OK) return rc;
    } else if (fisheye) {
        FBE_CUDA(cudaMemcpyAsync(v.out_kps_un, v.out_kps, (size_t)nimg * hplan.kp_cap_total * sizeof(fbe_keypoint), cudaMemcpyDeviceToDevice, stream));
    }
    if ((rc = launch_grid(hplan, dplan, v, nimg, stream)) != FBE_OK) return rc;
    FBE_MARK(5, stream);
    FBE_STAGE("grid", stream);
#undef FBE_STAGE
#undef FBE_MARK
    last_nimg = nimg;
    return FBE_OK;
}

int ExtractorCore::enable_timing(bool on) {
    FBE_CUDA(cudaSetDevice(cfg.device));
    if (on && tev.empty()) {
        tev.resize((size_t)kTimingRing * (kStages + 3));
        for (auto& e : tev) FBE_CUDA(cudaEventCreate(&e));
    }
    timing = on; tpos = 0; tcount = 0;
    return FBE_OK;
}

// Sums, over the recorded steps, the device time of: [0] pyramid, [1] FAST cells, [2] octree, [3] orient+describe,
// [4] grid, [5] blur (side stream).  Call after synchronising the stream.
int ExtractorCore::collect_timing(double* ms_sum, int* nsteps) {
    for (int i = 0; i < kStages; ++i) ms_sum[i] = 0.0;
    *nsteps = tcount;
    for (int k = 0; k < tcount; ++k) {
        cudaEvent_t* te = tev.data() + (size_t)k * (kStages + 3);
        float ms;
        const int a[kStages] = {0, 1, 2, 8, 4, 6}, b[kStages] = {1, 2, 3, 4, 5, 7};
        for (int i = 0; i < kStages; ++i) {
            FBE_CUDA(cudaEventElapsedTime(&ms, te[a[i]], te[b[i]]));
            ms_sum[i] += ms;
        }
    }
    return FBE_OK;
}

}  // namespace fbe
