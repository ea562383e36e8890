// Internal declarations shared by the sm_100a kernels and the C-ABI glue.  Not installed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>
#include <vector>
#include "../../include/fbe_cabi.h"

namespace fbe {

constexpr int kEdge = 19;        // EDGE_THRESHOLD (src/ORBextractor.cc:74): frame around every pyramid level
constexpr int kHalfPatch = 15;   // HALF_PATCH_SIZE (:73)
constexpr int kPatch = 31;       // PATCH_SIZE (:72)
constexpr int kMaxDim = 4095;    // x,y packed in 12 bits each
constexpr int kMaxGridCells = 64 * 48;
constexpr int kBlurTW = 128, kBlurTH = 64;   // blur output tile
constexpr int kRsTW = 128, kRsTH = 64;       // resize output tile (padded destination coordinates)
constexpr int kRsMaxTx = (kMaxDim + 38 + 15) / kRsTW + 2, kRsMaxTy = (kMaxDim + 38) / kRsTH + 2;
constexpr int kFastGroupW = 224; // widest run of FAST cells (pixels) one CTA of the FAST kernel owns

// packed candidate / key: score[31:24] | y[23:12] | x[11:0], level pixel coordinates
__host__ __device__ inline uint32_t pack_key(int x, int y, int s) { return ((uint32_t)s << 24) | ((uint32_t)y << 12) | (uint32_t)x; }
__host__ __device__ inline int key_x(uint32_t k) { return (int)(k & 0xFFFu); }
__host__ __device__ inline int key_y(uint32_t k) { return (int)((k >> 12) & 0xFFFu); }
__host__ __device__ inline int key_s(uint32_t k) { return (int)(k >> 24); }

// Geometry of one pyramid level for the current image size.  All offsets are per image slot.
struct LevelGeom {
    int w, h;            // level size (ROI)
    int pitch;           // bytes per padded row, multiple of 16
    int ph;              // padded rows = h + 38
    int img_off;         // byte offset of the padded level inside the per-slot pyramid slab (256-aligned)
    int ncols, nrows;    // FAST cell grid (src/ORBextractor.cc:781-787)
    int wcell, hcell;
    int wcell_recip;     // 65536 / wcell + 1: x / wcell == (x * wcell_recip) >> 16 for x < 256 (wcell in 30 .. 64)
    int cell_base;       // first cell of this level in the per-slot cell arrays
    int cell_cap;        // candidate slots per cell = ceil(wcell/2)*ceil(hcell/2) (NMS survivors are never adjacent)
    int gcells;          // FAST kernel: cells per CTA group (consecutive cells of one cell row, group width <= 256 px)
    int ngrp;            // groups per cell row
    int grp_base;        // first group of this level in the per-image group numbering
    int blur_ntx, blur_nty, blur_base;   // blur kernel: 128x32 output tiles in padded coordinates
    int rs_bw, rs_bh;    // resize kernel: TMA box (source tile) of a 128x64 output tile; rs_bw == 0 -> direct-global kernel
    int slot_base;       // first slot (u32 units) of this level in the per-slot slot array
    int key_cap;         // capacity of the compacted key array of this level ( = ncells*cell_cap )
    int nfeat;           // mnFeaturesPerLevel[level]
    int nini;            // octree roots
    float hx;            // root width
    int node_cap;        // live-node capacity of the octree
    int node_base;       // first node (in node-scratch units) of this level
    int kp_cap;          // selected keypoints capacity
    int kp_base;         // offset of this level in the per-slot selected-key array
    int tabx_off, taby_off;   // offsets (in entries) into the resize tables (levels >= 1)
    float scale;         // mvScaleFactor[level]
    float patch_size;    // (float)(int)(31*scale)
};

struct Plan {
    int nlevels;
    int rows, cols;
    int ini_th, min_th;
    int ncells_total;
    int ngroups_total;   // FAST CTAs per image
    int blur_tiles_total;
    int slots_total;
    int kp_cap_total;
    int nodes_total;
    int pyr_bytes;       // per-slot pyramid slab size
    int max_cell_w, max_cell_h;
    int umax[16];
    // grid-assignment parameters applied to the final (scaled) keypoints
    float grid_min_x, grid_min_y, grid_inv_w, grid_inv_h;
    int grid_cols, grid_rows;
    LevelGeom lv[FBE_MAX_LEVELS];
    // resize kernel: first source column (16-byte aligned) / row of every output tile column / tile row of a level
    short rs_x0[FBE_MAX_LEVELS][kRsMaxTx];
    short rs_y0[FBE_MAX_LEVELS][kRsMaxTy];
};

// Per-call device workspace pointers (all arrays are [max_batch] slabs, slot-major).
struct Workspace {
    int slot0;             // first batch slot of this view (TMA z coordinate = slot0 + image index)
    const uint8_t* in;     // [B][rows][in_pitch] source images
    int in_pitch, in_slot_stride;
    uint8_t* pyr;          // [B][pyr_bytes]
    uint8_t* blur;         // [B][pyr_bytes]  (same geometry as pyr; only the ROI is written)
    int* cell_count;       // [B][ncells_total]
    uint32_t* slots;       // [B][slots_total]
    uint32_t* keys;        // [B][slots_total]   compacted per level, level region = slot_base..
    uint32_t* key_node;    // [B][slots_total]
    uint8_t* oct_scratch;  // [B][oct_scratch_bytes]
    size_t oct_scratch_bytes;   // per slot
    uint32_t* sel;         // [B][kp_cap_total] selected keys per level in output order
    int* level_n;          // [B][FBE_MAX_LEVELS] keypoints per level
    fbe_keypoint* out_kps; // [B][kp_cap_total]   mvKeys
    fbe_keypoint* out_kps_un; // mvKeysUn: == out_kps unless a fisheye model is set (then its own [B][kp_cap_total] array)
    uint8_t* out_desc;     // [B][kp_cap_total][32]
    int* out_n;            // [B]
    int* out_cell;         // [B][kp_cap_total] grid cell id per keypoint or -1
    int* grid_start;       // [B][gcells+1]
    int* grid_items;       // [B][kp_cap_total]
    int* status;           // [B] error flags raised by kernels (capacity overflow ...)
    // per-plan CTA tables (one 32-bit load instead of a dependent walk over the levels):
    const uint32_t* fast_tab;   // [ngroups_total]   level << 24 | cell row << 12 | group in row
    const uint32_t* blur_tab;   // [blur_tiles_total] level << 24 | tile row << 12 | tile column
};

struct ResizeTab { int ofs; short a0, a1; };   // 8 bytes per padded coordinate

void set_error(const std::string& s);
const char* last_error();
extern std::atomic<unsigned long long> g_launches;
// Programmatic dependent launch: the extractor is a chain of short kernels, each consuming its predecessor's output.  A kernel
// launched with launch_dep() may start while its predecessor on the stream is still draining: CTAs are scheduled and run
// whatever precedes pdl_wait() (index arithmetic, barrier set-up, constant tables); pdl_wait() returns once the predecessor
// has completed and its writes are visible.  Launched without the attribute both calls are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... P, typename... A>
inline cudaError_t launch_dep(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t c = {};
    c.gridDim = grid; c.blockDim = block; c.dynamicSmemBytes = smem; c.stream = st;
    cudaLaunchAttribute at = {};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    c.attrs = &at; c.numAttrs = 1;
    return cudaLaunchKernelEx(&c, kern, P(std::forward<A>(args))...);
}
#endif

inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define FBE_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            ::fbe::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                        \
            return FBE_E_CUDA;                                                                           \
        }                                                                                                \
    } while (0)

// Per-thread device arena that only grows: a call carves its buffers out of one allocation, so the steady state issues no
// cudaMalloc / cudaFree (the entry point is re-entrant across threads like the other stateless calls of the C-ABI).
struct Arena {
    char* base = nullptr;
    size_t cap = 0, used = 0;
    int dev = -1;
    ~Arena() { if (base) cudaFree(base); }
    cudaError_t reserve(size_t bytes, int device) {
        used = 0;
        if (dev == device && cap >= bytes) return cudaSuccess;
        if (base) { cudaFree(base); base = nullptr; cap = 0; }
        dev = device;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&base), bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    template <class T> T* take(size_t count) {
        T* r = reinterpret_cast<T*>(base + used);
        used += (count * sizeof(T) + 255) & ~(size_t)255;
        return r;
    }
};
inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

struct TmaMaps;
// ---- kernel launchers (each enqueues on `st`) ---------------------------------------------------
int launch_pyramid(const Plan& hp, const Plan* dp, const Workspace& ws, const ResizeTab* d_tab, const TmaMaps& rs_maps, int nimg, cudaStream_t st);
int launch_fast_cells(const Plan& hp, const Plan* dp, const Workspace& ws, const TmaMaps& maps, int nimg, cudaStream_t st);
int launch_octree(const Plan& hp, const Plan* dp, const Workspace& ws, int nimg, cudaStream_t st);
int launch_blur(const Plan& hp, const Plan* dp, const Workspace& ws, const TmaMaps& maps, int nimg, cudaStream_t st);
int launch_describe(const Plan& hp, const Plan* dp, const Workspace& ws, int nimg, cudaStream_t st);
int launch_grid(const Plan& hp, const Plan* dp, const Workspace& ws, int nimg, cudaStream_t st);
int launch_undistort_batch(const fbe_keypoint* in, const int* n_arr, int stride, int nimg, const float K[4], const float D[4],
                           fbe_keypoint* out, cudaStream_t st);
size_t octree_scratch_bytes(const Plan& hp);
// stand-alone octree on caller candidates (device arrays): keys packed with level coordinates (= relative + 16)
int launch_octree_debug(const uint32_t* d_keys, uint32_t* d_knode, int nk, int nini, float hx, int H, int nfeat, int cap,
                        uint8_t* d_scratch, uint32_t* d_sel_idx, int* d_n, cudaStream_t st);
size_t octree_debug_scratch_bytes(int cap, int nini, int nfeat);

// stand-alone grid build on arbitrary keypoint arrays (device pointers), used by the matchers
int launch_grid_build(const fbe_keypoint* d_kps, const int* d_n, int n_stride, int nframes, float min_x, float min_y,
                      float inv_w, float inv_h, int gcols, int grows, int* d_cell_of, int* d_start, int* d_items,
                      cudaStream_t st);

// bird-view guidance + cornerSubPix on device-resident arrays (bird_refine.cu), chained by the bird feature block (bird_orb.cu)
int launch_bird_refine_dev(const uint8_t* d_contour, const uint8_t* d_img, int rows, int cols, int B, fbe_keypoint* d_in, const int* d_nin,
                           int cap, int half_w, int half_h, int max_iter, double eps, uint8_t* d_keep, fbe_keypoint* d_out, int* d_nkept,
                           int* d_iters, float* d_mask, fbe_keypoint** d_result, const int** d_result_n, cudaStream_t st);

// ---- extractor core: plan + workspace + run ------------------------------------------------------
struct ExtractorCore {
    fbe_extractor_cfg cfg;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> per_level;
    int umax[16];
    Plan hplan;            // valid when plan_rows/cols set
    Plan* dplan = nullptr;
    ResizeTab* dtab = nullptr;
    TmaMaps* blur_maps = nullptr;   // host copies of the per-level tensor maps over ws.pyr (passed by value at launch)
    TmaMaps* fast_maps = nullptr;   // box = 256 x (hcell + 6)
    TmaMaps* rs_maps = nullptr;     // m[l] = level l-1 as the SOURCE of level l, box = rs_bw x rs_bh
    int plan_rows = 0, plan_cols = 0;
    Workspace ws;
    uint8_t* d_in = nullptr;      // staging for host-API calls
    size_t d_in_bytes = 0;
    uint8_t* h_pin = nullptr;     // pinned staging
    size_t h_pin_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;   // blur runs beside FAST/octree
    cudaEvent_t ev_pyr = nullptr, ev_blur = nullptr;
    bool have_ws = false;
    int last_nimg = 0;
    // grid parameters (0 cols = grid disabled)
    float g_min_x = 0, g_min_y = 0, g_inv_w = 0, g_inv_h = 0;
    int g_cols = 0, g_rows = 0;

    // optional per-stage CUDA-event timing (bench.py's live roofline): ring of event sets, read back on demand
    static constexpr int kStages = 6;          // pyramid, fast, octree, describe, grid | blur (side stream)
    static constexpr int kTimingRing = 64;
    bool timing = false;
    std::vector<cudaEvent_t> tev;              // [ring][kStages + 3]
    int tpos = 0, tcount = 0;
    int enable_timing(bool on);
    int collect_timing(double* ms_sum /*[kStages]*/, int* nsteps);

    bool fisheye = false;          // Frame::UndistortKeyPoints on the device between describe and grid (set_fisheye)
    float fish_K[4] = {0, 0, 0, 0}, fish_D[4] = {0, 0, 0, 0};
    int set_fisheye(const float K[4], const float D[4]);
    int stream_priority = 0;       // CUDA priority of the main stream (0 = default, negative = more urgent); set before init
    int init(const fbe_extractor_cfg& c);
    void destroy();
    int ensure_plan(int rows, int cols);
    int build_workspace(int rows, int cols);   // the allocating part of ensure_plan
    int set_grid(float min_x, float min_y, float inv_w, float inv_h, int gcols, int grows);
    // images already on the device: [nimg][rows][pitch]; results go to workspace slots slot0 .. slot0+nimg-1
    // `out_set` selects one of `out_sets` copies of the OUTPUT arrays (keypoints, descriptors, counts, grid): the batch
    // pipeline alternates between two so that matching of step N can run beside the extraction of step N+1.
    int run_dev(const uint8_t* d_imgs, int pitch, int slot_stride, int nimg, int rows, int cols, int slot0 = 0, int out_set = 0);
    Workspace slot_view(int slot0, int out_set = 0) const;
    int out_sets = 1;      // set before the first ensure_plan()
    int free_ws();
};

void compute_extractor_tables(int nfeatures, float scale_factor, int nlevels, std::vector<float>& scale,
                              std::vector<float>& inv_scale, std::vector<float>& sigma2, std::vector<float>& inv_sigma2,
                              std::vector<int>& per_level, int umax[16]);
int build_plan(const fbe_extractor_cfg& cfg, const std::vector<float>& scale, const std::vector<float>& inv_scale,
               const std::vector<int>& per_level, const int umax[16], int rows, int cols, Plan& p,
               std::vector<ResizeTab>& tabs);

}  // namespace fbe
