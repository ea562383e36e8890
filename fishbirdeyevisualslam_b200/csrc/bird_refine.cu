// The per-keypoint steps of the reference's bird-view feature path between detection and description:
//   Frame::GuidenceKeyBirdPts / nearEdges  (src/Frame.cc:671-684, 717-739): keep a keypoint when any pixel of the ~21x21 window
//     of mBirdviewContourICP around it is >= 10 (edge or free space), in input order;
//   cv::cornerSubPix(mBirdviewImg, pts, Size(5,5), Size(-1,-1), {EPS+MAX_ITER, 40, 0.001})  (src/Frame.cc:349-352).
// Both are one warp per keypoint.  nearEdges: the lanes stride over the window, one ballot per 32 pixels, early exit.
// cornerSubPix: per iteration the lanes sample the (2*hw+3) x (2*hh+3) bilinear patch into shared memory (cv::getRectSubPix
// arithmetic), compute the five per-pixel terms of the normal equations in double, and five lanes add one accumulator each IN
// PIXEL ORDER -- the reference's sequential double sums, so the iteration count and the float result are those of the CPU code
// (pinned to cv2 4.13.0 for windows inside the image; tests/test_bird_refine.py).  The exp() weights come from the host so that
// they are glibc's, like the reference's.
#include <cfloat>
#include <cmath>
#include <vector>
#include "fbe_internal.cuh"

namespace fbe {

constexpr int kWarpsPerCta = 4;

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_near_edges(const uint8_t* __restrict__ contour, int rows, int cols, size_t step, size_t img_stride, const fbe_keypoint* __restrict__ kps,
             const int* __restrict__ n_arr, int cap, uint8_t* __restrict__ keep) {
    const int b = blockIdx.y;                      // frame of the batch: images img_stride bytes apart, keypoint lists cap records apart
    contour += (size_t)b * img_stride; kps += (size_t)b * cap; keep += (size_t)b * cap;
    const int n = min(n_arr[b], cap);
    const int k = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n) return;
    const float r = 10.f, x = kps[k].x, y = kps[k].y;
    const float pt1x = (x - r) > 0 ? (x - r) : 0.f, pt1y = (y - r) > 0 ? (y - r) : 0.f;
    const float pt2x = (x + r) < (float)cols ? (x + r) : (float)cols, pt2y = (y + r) < (float)rows ? (y + r) : (float)rows;
    bool hit = false;
    if (pt1x < pt2x && pt1y < pt2y) {
        // `size_t row = pt1x; row < pt2x` : truncation below, first integer not below pt2x above
        const int r0 = (int)pt1x, r1 = (int)ceilf(pt2x), c0 = (int)pt1y, c1 = (int)ceilf(pt2y);
        const int nc = c1 - c0, total = (r1 - r0) * nc;
        const size_t limit = (size_t)rows * step;
        for (int base = 0; base < total; base += 32) {
            const int i = base + lane;
            bool h = false;
            if (i < total) {
                const size_t a = (size_t)(r0 + i / nc) * step + (size_t)(c0 + i % nc);   // at<uchar>(row = x range, col = y range)
                h = a < limit && contour[a] >= 10;
            }
            if (__ballot_sync(0xffffffffu, h)) { hit = true; break; }
        }
    }
    if (lane == 0) keep[k] = hit ? 1 : 0;
}

// ordered compaction of the kept keypoints (mvKeysBird.push_back in input order); n is a few thousand: one CTA
__global__ void __launch_bounds__(1024)
k_compact_kept(const fbe_keypoint* __restrict__ kps, const uint8_t* __restrict__ keep, const int* __restrict__ n_arr, int cap,
               fbe_keypoint* __restrict__ out, int* __restrict__ n_out) {
    const int b = blockIdx.y, n = min(n_arr[b], cap);
    kps += (size_t)b * cap; keep += (size_t)b * cap; out += (size_t)b * cap; n_out += b;
    __shared__ int warp_sum[32];
    __shared__ int running;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const bool kp = i < n && keep[i];
        const unsigned m = __ballot_sync(0xffffffffu, kp);
        if (lane == 0) warp_sum[w] = __popc(m);
        __syncthreads();
        int before = running;
        for (int j = 0; j < w; ++j) before += warp_sum[j];
        if (kp) out[before + __popc(m & ((1u << lane) - 1u))] = kps[i];
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int j = 0; j < 32; ++j) t += warp_sum[j]; running += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = running;
}

// one bilinear sample of cv::getRectSubPix(8U -> 32F): window origin (ipx, ipy), fractions folded into a11..b2
__device__ __forceinline__ float rect_sample(const uint8_t* __restrict__ img, int rows, int cols, size_t step, int ipx, int ipy, int i, int j,
                                             bool inside, int rx, int rw, float a11, float a12, float a21, float a22, float b1, float b2) {
    if (inside) {
        const uint8_t* p = img + (size_t)(ipy + i) * step + (ipx + j);
        return ((float)p[0] * a11 + (float)p[1] * a12) + ((float)p[step] * a21 + (float)p[step + 1] * a22);
    }
    // window crosses the border: rows clamp (replicated), columns left of rx / right of rw take the edge column
    const int yt = min(max(ipy + i, 0), rows - 1), yb = min(max(ipy + i + 1, 0), rows - 1);
    const uint8_t* pt = img + (size_t)yt * step;
    const uint8_t* pb = img + (size_t)yb * step;
    if (j < rx || j >= rw) {
        const int c = j < rx ? max(ipx, 0) : cols - 1;
        return (float)pt[c] * b1 + (float)pb[c] * b2;
    }
    const int c = ipx + j;
    return ((float)pt[c] * a11 + (float)pt[c + 1] * a12) + ((float)pb[c] * a21 + (float)pb[c + 1] * a22);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_corner_subpix(const uint8_t* __restrict__ img, int rows, int cols, size_t step, size_t img_stride, fbe_keypoint* __restrict__ kps,
                const int* __restrict__ n_arr, int cap, int hw, int hh, int max_iter, double eps2, const float* __restrict__ mask,
                int* __restrict__ iters) {
    extern __shared__ double smem_d[];
    const int b = blockIdx.y, n = min(n_arr[b], cap);
    img += (size_t)b * img_stride; kps += (size_t)b * cap; iters += (size_t)b * cap;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * kWarpsPerCta + w;
    if (k >= n) return;
    const int ww = 2 * hw + 1, wh = 2 * hh + 1, pw = ww + 2, ph = wh + 2, nq = ww * wh;
    double* terms = smem_d + (size_t)w * (5 * nq + (pw * ph + 1) / 2);       // [5][nq] doubles, then the float patch
    float* patch = reinterpret_cast<float*>(terms + 5 * nq);

    // (e + 0.5) / pw is at least 0.5 / pw away from an integer: the float product truncates to the exact quotient for these sizes
    const float inv_pw = 1.f / (float)pw, inv_ww = 1.f / (float)ww;
    const float tx = kps[k].x, ty = kps[k].y;
    float ix = tx, iy = ty;
    int iter = 0;
    double err = 0;
    do {
        // ---- getRectSubPix(img, (pw, ph), (ix, iy)) -> patch
        const float ox = ix - (pw - 1) * 0.5f, oy = iy - (ph - 1) * 0.5f;
        const int ipx = (int)floorf(ox), ipy = (int)floorf(oy);
        const float a = ox - (float)ipx, b = oy - (float)ipy;
        const float a11 = (1.f - a) * (1.f - b), a12 = a * (1.f - b), a21 = (1.f - a) * b, a22 = a * b, b1 = 1.f - b, b2 = b;
        const bool inside = 0 <= ipx && ipx < cols - pw && 0 <= ipy && ipy < rows - ph;
        int rx = 0, rw = pw;
        if (!inside) {
            if (ipx < 0) rx = min(-ipx, pw);
            if (!(ipx < cols - pw)) rw = max(cols - ipx - 1, 0);
        }
        for (int e = lane; e < pw * ph; e += 32) {
            const int i = (int)(((float)e + 0.5f) * inv_pw), j = e - i * pw;       // e / pw, e % pw without the emulated division
            patch[e] = rect_sample(img, rows, cols, step, ipx, ipy, i, j, inside, rx, rw, a11, a12, a21, a22, b1, b2);
        }
        __syncwarp();
        // ---- per-pixel terms of the 2x2 normal equations (double, like the reference)
        for (int q = lane; q < nq; q += 32) {
            const int i = (int)(((float)q + 0.5f) * inv_ww), j = q - i * ww;
            const float* sp = patch + (i + 1) * pw + (j + 1);
            const double m = (double)mask[q];
            const double tgx = (double)(sp[1] - sp[-1]);
            const double tgy = (double)(sp[pw] - sp[-pw]);
            const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
            const double px = (double)(j - hw), py = (double)(i - hh);
            terms[q] = gxx; terms[nq + q] = gxy; terms[2 * nq + q] = gyy;
            terms[3 * nq + q] = gxx * px + gxy * py;
            terms[4 * nq + q] = gxy * px + gyy * py;
        }
        __syncwarp();
        // ---- five lanes add one accumulator each, in pixel order (sequential sums of the reference)
        double acc = 0;
        if (lane < 5) {
            const double* t = terms + lane * nq;
            // the adds form one dependent chain (that IS the reference order); unrolling lets the shared-memory loads of the
            // next terms issue ahead of it, so the chain runs at the DADD latency instead of load + add
#pragma unroll 1
            for (int q0 = 0; q0 + 11 <= nq; q0 += 11) {
                double v[11];
#pragma unroll
                for (int u = 0; u < 11; ++u) v[u] = t[q0 + u];
#pragma unroll
                for (int u = 0; u < 11; ++u) acc += v[u];
            }
            for (int q = nq - nq % 11; q < nq; ++q) acc += t[q];
        }
        const double sa = __shfl_sync(0xffffffffu, acc, 0), sb = __shfl_sync(0xffffffffu, acc, 1), sc = __shfl_sync(0xffffffffu, acc, 2);
        const double bb1 = __shfl_sync(0xffffffffu, acc, 3), bb2 = __shfl_sync(0xffffffffu, acc, 4);
        const double det = sa * sc - sb * sb;
        if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
        const double scale = 1.0 / det;
        const float nx = (float)((double)ix + sc * scale * bb1 - sb * scale * bb2);
        const float ny = (float)((double)iy - sb * scale * bb1 + sa * scale * bb2);
        err = (double)(nx - ix) * (double)(nx - ix) + (double)(ny - iy) * (double)(ny - iy);
        ix = nx; iy = ny;
        if (ix < 0 || ix >= (float)cols || iy < 0 || iy >= (float)rows) { ++iter; break; }
    } while (++iter < max_iter && err > eps2);
    if (fabsf(ix - tx) > (float)hw || fabsf(iy - ty) > (float)hh) { ix = tx; iy = ty; }   // moved too far: keep the input point
    if (lane == 0) {
        kps[k].x = ix; kps[k].y = iy;
        iters[k] = iter;
    }
}

static inline bool img_too_large(int hw, int hh) { return hw < 1 || hh < 1 || hw > 10 || hh > 10; }

static void subpix_mask(int hw, int hh, std::vector<float>& mask) {
    const int ww = 2 * hw + 1, wh = 2 * hh + 1;
    mask.resize((size_t)ww * wh);
    for (int i = 0; i < wh; ++i) {
        const float y = (float)(i - hh) / hh;
        const float vy = std::exp(-y * y);
        for (int j = 0; j < ww; ++j) {
            const float x = (float)(j - hw) / hw;
            mask[(size_t)i * ww + j] = (float)(vy * std::exp(-x * x));
        }
    }
}

// Device-resident form for the bird feature block (bird_orb.cu chains detect -> this -> compute without a host round trip):
// contour / img are [B][rows][cols] tightly packed device images (either may be NULL), d_in [B][cap] keypoints with counts
// d_nin[B].  With a contour the kept keypoints go to d_out / d_nkept (ordered); cornerSubPix then refines that list in place.
// *d_result / *d_result_n name the arrays that hold the final list.  d_mask: 441 floats of scratch.
int launch_bird_refine_dev(const uint8_t* d_contour, const uint8_t* d_img, int rows, int cols, int B, fbe_keypoint* d_in, const int* d_nin,
                           int cap, int half_w, int half_h, int max_iter, double eps, uint8_t* d_keep, fbe_keypoint* d_out, int* d_nkept,
                           int* d_iters, float* d_mask, fbe_keypoint** d_result, const int** d_result_n, cudaStream_t st) {
    if (img_too_large(half_w, half_h)) return FBE_E_INVALID;
    const size_t total = (size_t)B * cap, img_bytes = (size_t)rows * cols;
    const dim3 grid((cap + kWarpsPerCta - 1) / kWarpsPerCta, B);
    fbe_keypoint* d_cur = d_in;
    const int* d_count = d_nin;
    if (d_contour) {
        FBE_CUDA(cudaMemsetAsync(d_keep, 0, total, st));
        k_near_edges<<<grid, kWarpsPerCta * 32, 0, st>>>(d_contour, rows, cols, (size_t)cols, img_bytes, d_cur, d_nin, cap, d_keep);
        k_compact_kept<<<dim3(1, B), 1024, 0, st>>>(d_cur, d_keep, d_nin, cap, d_out, d_nkept);
        count_launch(2);
        d_cur = d_out;
        d_count = d_nkept;
    }
    if (d_img) {
        std::vector<float> mask;
        subpix_mask(half_w, half_h, mask);
        FBE_CUDA(cudaMemcpyAsync(d_mask, mask.data(), mask.size() * sizeof(float), cudaMemcpyHostToDevice, st));   // pageable: staged before return
        FBE_CUDA(cudaMemsetAsync(d_iters, 0, total * sizeof(int), st));
        if (max_iter < 1) max_iter = 1;
        if (max_iter > 100) max_iter = 100;
        double e2 = eps > 0 ? eps : 0;
        e2 *= e2;
        const int nq = (2 * half_w + 1) * (2 * half_h + 1), np = (2 * half_w + 3) * (2 * half_h + 3);
        const size_t smem = (size_t)kWarpsPerCta * (5 * nq + (np + 1) / 2) * sizeof(double);
        FBE_CUDA(cudaFuncSetAttribute(k_corner_subpix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_corner_subpix<<<grid, kWarpsPerCta * 32, smem, st>>>(d_img, rows, cols, (size_t)cols, img_bytes, d_cur, d_count, cap, half_w, half_h,
                                                               max_iter, e2, d_mask, d_iters);
        count_launch();
    }
    FBE_CUDA(cudaGetLastError());
    *d_result = d_cur;
    *d_result_n = d_count;
    return FBE_OK;
}

}  // namespace fbe

using namespace fbe;

// B frames of one size: contours / imgs are B images `*_stride` bytes apart (rows of `*_step` bytes), kps / out_kps / keep / iters
// B lists `cap` records apart, n / n_out B counts.  The single-frame entry point is the B = 1 case.
static int bird_refine_impl(const uint8_t* contour, size_t contour_step, size_t contour_stride, const uint8_t* img, size_t img_step,
                            size_t img_stride, int rows, int cols, int B, const fbe_keypoint* kps, const int32_t* n, int cap, int half_w,
                            int half_h, int max_iter, double eps, int device, uint8_t* keep, fbe_keypoint* out_kps, int32_t* n_out,
                            int32_t* iters) {
    if (B <= 0 || cap < 0 || rows <= 0 || cols <= 0 || (!contour && !img) || !n || !n_out || (cap > 0 && (!kps || !out_kps))) return FBE_E_INVALID;
    if (contour && contour_step < (size_t)cols) return FBE_E_INVALID;
    if (img && (img_step < (size_t)cols || half_w < 1 || half_h < 1 || half_w > 10 || half_h > 10)) return FBE_E_INVALID;
    int n_max = 0;
    for (int b = 0; b < B; ++b) {
        if (n[b] < 0 || n[b] > cap) return FBE_E_INVALID;
        n_max = n[b] > n_max ? n[b] : n_max;
        n_out[b] = 0;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device: this library has no CPU path"); return FBE_E_CUDA; }
    FBE_CUDA(cudaSetDevice(device));
    if (n_max == 0) return FBE_OK;
    static thread_local Arena arena;
    const size_t total = (size_t)B * cap, kp_bytes = total * sizeof(fbe_keypoint), img_bytes = (size_t)rows * cols;
    const int nq = (2 * half_w + 1) * (2 * half_h + 1), np = (2 * half_w + 3) * (2 * half_h + 3);
    FBE_CUDA(arena.reserve(2 * pad256(kp_bytes) + pad256(total) + 2 * pad256((size_t)B * sizeof(int)) + pad256(total * sizeof(int)) +
                          2 * pad256(img_bytes * B) + pad256(441 * sizeof(float)), device));
    fbe_keypoint* d_in = arena.take<fbe_keypoint>(total);
    fbe_keypoint* d_out = arena.take<fbe_keypoint>(total);
    uint8_t* d_keep = arena.take<uint8_t>(total);
    int* d_nin = arena.take<int>(B);
    int* d_nkept = arena.take<int>(B);
    int* d_iters = arena.take<int>(total);
    uint8_t* d_contour = arena.take<uint8_t>(img_bytes * B);
    uint8_t* d_img = arena.take<uint8_t>(img_bytes * B);
    float* d_mask = arena.take<float>(441);
    cudaStream_t st = cudaStreamPerThread;
    FBE_CUDA(cudaMemcpyAsync(d_in, kps, kp_bytes, cudaMemcpyHostToDevice, st));
    FBE_CUDA(cudaMemcpyAsync(d_nin, n, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, st));
    const dim3 grid((n_max + kWarpsPerCta - 1) / kWarpsPerCta, B);
    fbe_keypoint* d_cur = d_in;
    const int* d_count = d_nin;
    auto upload = [&](uint8_t* dst, const uint8_t* src, size_t step, size_t stride) -> cudaError_t {
        if (stride == step * (size_t)rows || B == 1)       // frames back to back: one 2-D copy over B*rows rows
            return cudaMemcpy2DAsync(dst, (size_t)cols, src, step, (size_t)cols, (size_t)rows * B, cudaMemcpyHostToDevice, st);
        for (int b = 0; b < B; ++b) {
            cudaError_t e = cudaMemcpy2DAsync(dst + img_bytes * b, (size_t)cols, src + stride * b, step, (size_t)cols, (size_t)rows, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    if (contour) {          // GuidenceKeyBirdPts: filter + ordered compaction
        FBE_CUDA(upload(d_contour, contour, contour_step, contour_stride));
        FBE_CUDA(cudaMemsetAsync(d_keep, 0, total, st));          // entries past n[f] of a ragged batch read as not kept
        k_near_edges<<<grid, kWarpsPerCta * 32, 0, st>>>(d_contour, rows, cols, (size_t)cols, img_bytes, d_cur, d_nin, cap, d_keep);
        k_compact_kept<<<dim3(1, B), 1024, 0, st>>>(d_cur, d_keep, d_nin, cap, d_out, d_nkept);
        count_launch(2);
        d_cur = d_out;
        d_count = d_nkept;
    }
    if (img) {              // cornerSubPix on the kept points (counts read on the device: no host round trip in between)
        FBE_CUDA(upload(d_img, img, img_step, img_stride));
        std::vector<float> mask;
        subpix_mask(half_w, half_h, mask);
        FBE_CUDA(cudaMemcpyAsync(d_mask, mask.data(), mask.size() * sizeof(float), cudaMemcpyHostToDevice, st));   // pageable: staged before return
        FBE_CUDA(cudaMemsetAsync(d_iters, 0, total * sizeof(int), st));
        if (max_iter < 1) max_iter = 1;
        if (max_iter > 100) max_iter = 100;
        double e2 = eps > 0 ? eps : 0;
        e2 *= e2;
        const size_t smem = (size_t)kWarpsPerCta * (5 * nq + (np + 1) / 2) * sizeof(double);
        FBE_CUDA(cudaFuncSetAttribute(k_corner_subpix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_corner_subpix<<<grid, kWarpsPerCta * 32, smem, st>>>(d_img, rows, cols, (size_t)cols, img_bytes, d_cur, d_count, cap, half_w, half_h,
                                                               max_iter, e2, d_mask, d_iters);
        count_launch();
    }
    FBE_CUDA(cudaGetLastError());
    FBE_CUDA(cudaMemcpyAsync(n_out, d_count, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (keep) {
        if (contour) FBE_CUDA(cudaMemcpyAsync(keep, d_keep, total, cudaMemcpyDeviceToHost, st));
        else for (int b = 0; b < B; ++b) for (int i = 0; i < cap; ++i) keep[(size_t)b * cap + i] = i < n[b];
    }
    FBE_CUDA(cudaMemcpyAsync(out_kps, d_cur, kp_bytes, cudaMemcpyDeviceToHost, st));
    if (iters && img) FBE_CUDA(cudaMemcpyAsync(iters, d_iters, total * sizeof(int), cudaMemcpyDeviceToHost, st));
    FBE_CUDA(cudaStreamSynchronize(st));
    return FBE_OK;
}

extern "C" {

int fbe_bird_refine(const uint8_t* contour, size_t contour_step, const uint8_t* img, size_t img_step, int32_t rows, int32_t cols,
                    const fbe_keypoint* kps, int32_t n, int32_t half_w, int32_t half_h, int32_t max_iter, double eps, int32_t device,
                    uint8_t* keep, fbe_keypoint* out_kps, int32_t* n_out, int32_t* iters) {
    if (n < 0 || !n_out) return FBE_E_INVALID;
    return bird_refine_impl(contour, contour_step, 0, img, img_step, 0, rows, cols, 1, kps, &n, n, half_w, half_h, max_iter, eps, device, keep,
                            out_kps, n_out, iters);
}

int fbe_bird_refine_batch(const uint8_t* contours, size_t contour_step, size_t contour_stride, const uint8_t* imgs, size_t img_step,
                          size_t img_stride, int32_t rows, int32_t cols, int32_t nframes, const fbe_keypoint* kps, const int32_t* n,
                          int32_t cap, int32_t half_w, int32_t half_h, int32_t max_iter, double eps, int32_t device, uint8_t* keep,
                          fbe_keypoint* out_kps, int32_t* n_out, int32_t* iters) {
    return bird_refine_impl(contours, contour_step, contour_stride, imgs, img_step, img_stride, rows, cols, nframes, kps, n, cap, half_w, half_h,
                            max_iter, eps, device, keep, out_kps, n_out, iters);
}

}  // extern "C"
