// ORACLE (test infrastructure): C entry points around the reference's OWN ORBextractor, compiled
// verbatim from /root/reference/src/ORBextractor.cc against oracle/cvshim (see oracle/Makefile).
//
// A monotonic, per-thread bump arena backs operator new while an extraction runs, so that
// `ExtractorNode*` order == creation order.  That canonicalises the one address-dependent step in
// the reference (std::sort over pair<int, ExtractorNode*>, src/ORBextractor.cc:684).
#include <sys/mman.h>
#include <cstdio>
#include <cstdlib>
#include <new>
#include "opencv2/core/core.hpp"
#include "ORBextractor.h"

namespace {
struct Arena { char* base; size_t cap; size_t off; bool active; };
thread_local Arena g_arena = {nullptr, 0, 0, false};
const size_t kArenaBytes = (size_t)8 << 30;   // virtual reservation, touched lazily

void arena_begin() {
    if (!g_arena.base) {
        void* p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { std::perror("oracle arena mmap"); std::abort(); }
        g_arena.base = (char*)p; g_arena.cap = kArenaBytes;
    }
    g_arena.off = 0; g_arena.active = true;
}
void arena_end() { g_arena.active = false; }
inline bool in_arena(void* p) { return g_arena.base && (char*)p >= g_arena.base && (char*)p < g_arena.base + g_arena.cap; }
}  // namespace

void* operator new(size_t n) {
    if (g_arena.active) {
        size_t a = (g_arena.off + 15) & ~(size_t)15;
        if (a + n > g_arena.cap) { std::fprintf(stderr, "oracle arena exhausted\n"); std::abort(); }
        g_arena.off = a + n;
        return g_arena.base + a;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { if (p && !in_arena(p)) std::free(p); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

extern "C" {

void* ref_extractor_create(int nfeatures, float scale, int nlevels, int ini_th, int min_th) {
    return new ORB_SLAM2::ORBextractor(nfeatures, scale, nlevels, ini_th, min_th);
}
void ref_extractor_destroy(void* h) { delete (ORB_SLAM2::ORBextractor*)h; }

// Scale tables (GetScaleFactors etc.), each nlevels floats.
void ref_extractor_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(), d = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < e->GetLevels(); ++i) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
}

// Runs operator() on an 8-bit image.  kps: capacity x 7 floats/ints in cv::KeyPoint layout (28 B).
// Returns the number of keypoints (may exceed capacity; then only `capacity` are written).
int ref_extract(void* h, const uint8_t* img, int rows, int cols, int step, void* kps, uint8_t* desc, int capacity) {
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    int n = 0;
    arena_begin();
    {
        cv::Mat im(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; ++y) std::memcpy(im.ptr(y), img + (size_t)y * step, cols);
        std::vector<cv::KeyPoint> keys;
        cv::Mat d;
        (*e)(im, cv::Mat(), keys, d);
        n = (int)keys.size();
        int m = std::min(n, capacity);
        if (m > 0) {
            std::memcpy(kps, keys.data(), (size_t)m * sizeof(cv::KeyPoint));
            for (int i = 0; i < m; ++i) std::memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
        }
        // The pyramid must not outlive the arena: drop the views before the arena is rewound.
        for (size_t l = 0; l < e->mvImagePyramid.size(); ++l) e->mvImagePyramid[l] = cv::Mat();
    }
    arena_end();
    return n;
}

// Same, but also copies out pyramid level `level` WITH its 19-px border (for the pyramid parity test).
int ref_pyramid_level(void* h, const uint8_t* img, int rows, int cols, int step, int level,
                      uint8_t* dst, int dst_step, int* out_rows, int* out_cols) {
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    arena_begin();
    {
        cv::Mat im(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; ++y) std::memcpy(im.ptr(y), img + (size_t)y * step, cols);
        std::vector<cv::KeyPoint> keys;
        cv::Mat d;
        (*e)(im, cv::Mat(), keys, d);
        const cv::Mat& L = e->mvImagePyramid[level];
        *out_rows = L.rows; *out_cols = L.cols;
        if (dst) {
            const uint8_t* base = L.data - 19 * L.step - 19;
            for (int y = 0; y < L.rows + 38; ++y) std::memcpy(dst + (size_t)y * dst_step, base + (size_t)y * L.step, L.cols + 38);
        }
        for (size_t l = 0; l < e->mvImagePyramid.size(); ++l) e->mvImagePyramid[l] = cv::Mat();
    }
    arena_end();
    return 0;
}

int ref_sizeof_keypoint() { return (int)sizeof(cv::KeyPoint); }

}  // extern "C"
