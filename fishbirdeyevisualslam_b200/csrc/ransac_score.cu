// Initializer::CheckHomography / CheckFundamental (src/Initializer.cc:391-554) for all RANSAC hypotheses of one
// FindHomography / FindFundamental call: K hypotheses x N matches symmetric transfer / epipolar errors.
// One CTA per hypothesis: its threads evaluate the per-match terms in parallel (each term is the reference's float
// expression, operation by operation, no FMA; `1.0/(...)` is a double division rounded to float as written), then ONE
// thread adds the terms in match order -- float addition is not associative and the reference accumulates sequentially, so
// this is what makes the score bit-identical.  Terms live in shared memory in chunks; a rejected term is skipped, not
// added as zero (adding +0.0f would be harmless, but the skip mirrors the reference's control flow).
#include <cmath>
#include "fbe_internal.cuh"

namespace fbe {

constexpr int kScoreThreads = 256;
constexpr int kScoreChunk = 2048;          // matches per shared-memory chunk (2 terms each)

struct ScoreIn {
    const fbe_keypoint* kps1; const fbe_keypoint* kps2; const int2* matches; int n;
    const float* A;        // K x 9: H21 or F21
    const float* B;        // K x 9: H12 (homography only)
    float inv_sigma2;
    float* scores; uint8_t* inliers;
};

template <bool kHomography>
__global__ void __launch_bounds__(kScoreThreads) k_ransac_score(const ScoreIn in) {
    __shared__ float s_term[2 * kScoreChunk];
    __shared__ uint8_t s_ok[2 * kScoreChunk];
    __shared__ float s_score;
    const int k = blockIdx.x, tid = threadIdx.x;
    const float* a = in.A + (size_t)k * 9;
    const float a11 = a[0], a12 = a[1], a13 = a[2], a21 = a[3], a22 = a[4], a23 = a[5], a31 = a[6], a32 = a[7], a33 = a[8];
    float b11 = 0, b12 = 0, b13 = 0, b21 = 0, b22 = 0, b23 = 0, b31 = 0, b32 = 0, b33 = 0;
    if (kHomography) {
        const float* b = in.B + (size_t)k * 9;
        b11 = b[0]; b12 = b[1]; b13 = b[2]; b21 = b[3]; b22 = b[4]; b23 = b[5]; b31 = b[6]; b32 = b[7]; b33 = b[8];
    }
    const float th = kHomography ? 5.991f : 3.841f, th_score = 5.991f;
    if (tid == 0) s_score = 0.0f;
    for (int base = 0; base < in.n; base += kScoreChunk) {
        const int cnt = min(kScoreChunk, in.n - base);
        __syncthreads();
        for (int j = tid; j < cnt; j += kScoreThreads) {
            const int2 mt = in.matches[base + j];
            const float u1 = in.kps1[mt.x].x, v1 = in.kps1[mt.x].y, u2 = in.kps2[mt.y].x, v2 = in.kps2[mt.y].y;
            float chi1, chi2;
            if (kHomography) {
                // x2in1 = H12*x2 (:437-445)
                const float w2 = (float)(1.0 / (double)__fadd_rn(__fadd_rn(__fmul_rn(b31, u2), __fmul_rn(b32, v2)), b33));
                const float u2in1 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(b11, u2), __fmul_rn(b12, v2)), b13), w2);
                const float v2in1 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(b21, u2), __fmul_rn(b22, v2)), b23), w2);
                const float du1 = __fsub_rn(u1, u2in1), dv1 = __fsub_rn(v1, v2in1);
                chi1 = __fmul_rn(__fadd_rn(__fmul_rn(du1, du1), __fmul_rn(dv1, dv1)), in.inv_sigma2);
                // x1in2 = H21*x1 (:455-463)
                const float w1 = (float)(1.0 / (double)__fadd_rn(__fadd_rn(__fmul_rn(a31, u1), __fmul_rn(a32, v1)), a33));
                const float u1in2 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(a11, u1), __fmul_rn(a12, v1)), a13), w1);
                const float v1in2 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(a21, u1), __fmul_rn(a22, v1)), a23), w1);
                const float du2 = __fsub_rn(u2, u1in2), dv2 = __fsub_rn(v2, v1in2);
                chi2 = __fmul_rn(__fadd_rn(__fmul_rn(du2, du2), __fmul_rn(dv2, dv2)), in.inv_sigma2);
            } else {
                // l2 = F21 x1 (:514-522)
                const float a2 = __fadd_rn(__fadd_rn(__fmul_rn(a11, u1), __fmul_rn(a12, v1)), a13);
                const float b2 = __fadd_rn(__fadd_rn(__fmul_rn(a21, u1), __fmul_rn(a22, v1)), a23);
                const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(a31, u1), __fmul_rn(a32, v1)), a33);
                const float num2 = __fadd_rn(__fadd_rn(__fmul_rn(a2, u2), __fmul_rn(b2, v2)), c2);
                chi1 = __fmul_rn(__fdiv_rn(__fmul_rn(num2, num2), __fadd_rn(__fmul_rn(a2, a2), __fmul_rn(b2, b2))), in.inv_sigma2);
                // l1 = x2' F21 (:534-542)
                const float a1 = __fadd_rn(__fadd_rn(__fmul_rn(a11, u2), __fmul_rn(a21, v2)), a31);
                const float b1 = __fadd_rn(__fadd_rn(__fmul_rn(a12, u2), __fmul_rn(a22, v2)), a32);
                const float c1 = __fadd_rn(__fadd_rn(__fmul_rn(a13, u2), __fmul_rn(a23, v2)), a33);
                const float num1 = __fadd_rn(__fadd_rn(__fmul_rn(a1, u1), __fmul_rn(b1, v1)), c1);
                chi2 = __fmul_rn(__fdiv_rn(__fmul_rn(num1, num1), __fadd_rn(__fmul_rn(a1, a1), __fmul_rn(b1, b1))), in.inv_sigma2);
            }
            const bool ok1 = !(chi1 > th), ok2 = !(chi2 > th);               // NaN: `chi > th` is false -> counted, like the reference
            s_term[2 * j] = __fsub_rn(th_score, chi1); s_ok[2 * j] = ok1;
            s_term[2 * j + 1] = __fsub_rn(th_score, chi2); s_ok[2 * j + 1] = ok2;
            if (in.inliers) in.inliers[(size_t)k * in.n + base + j] = ok1 && ok2;
        }
        __syncthreads();
        if (tid == 0) {
            float score = s_score;
            for (int j = 0; j < 2 * cnt; ++j)
                if (s_ok[j]) score = __fadd_rn(score, s_term[j]);
            s_score = score;
        }
    }
    __syncthreads();
    if (tid == 0) in.scores[k] = s_score;
}

static int run_score(bool homography, const fbe_keypoint* kps1, const fbe_keypoint* kps2, const int32_t* matches, int32_t n,
                     const float* A, const float* B, int32_t K, float sigma, int32_t device, float* scores, uint8_t* inliers) {
    if (n < 0 || K < 0 || (K > 0 && (!A || !scores || (homography && !B))) || (n > 0 && (!kps1 || !kps2 || !matches))) return FBE_E_INVALID;
    if (K == 0) return FBE_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device: this library has no CPU path"); return FBE_E_CUDA; }
    FBE_CUDA(cudaSetDevice(device));
    int max1 = -1, max2 = -1;
    for (int i = 0; i < n; ++i) {
        if (matches[2 * i] < 0 || matches[2 * i + 1] < 0) return FBE_E_INVALID;
        max1 = std::max(max1, matches[2 * i]); max2 = std::max(max2, matches[2 * i + 1]);
    }
    const size_t b_k1 = (size_t)(max1 + 1) * sizeof(fbe_keypoint), b_k2 = (size_t)(max2 + 1) * sizeof(fbe_keypoint);
    const size_t b_m = (size_t)n * 8, b_A = (size_t)K * 36, b_s = (size_t)K * 4, b_in = inliers ? (size_t)K * n : 0;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t total = up(b_k1) + up(b_k2) + up(b_m) + 2 * up(b_A) + up(b_s) + up(b_in) + 256;
    uint8_t* d = nullptr;
    FBE_CUDA(cudaMalloc(&d, total));
    uint8_t* p = d;
    auto take = [&](size_t bytes) { uint8_t* r = p; p += up(bytes); return r; };
    fbe_keypoint* d_k1 = (fbe_keypoint*)take(b_k1); fbe_keypoint* d_k2 = (fbe_keypoint*)take(b_k2);
    int2* d_m = (int2*)take(b_m); float* d_A = (float*)take(b_A); float* d_B = (float*)take(b_A);
    float* d_s = (float*)take(b_s); uint8_t* d_in = inliers ? take(b_in) : nullptr;
    cudaError_t e = cudaSuccess;
    if (b_k1) e = cudaMemcpy(d_k1, kps1, b_k1, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && b_k2) e = cudaMemcpy(d_k2, kps2, b_k2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && b_m) e = cudaMemcpy(d_m, matches, b_m, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_A, A, b_A, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && homography) e = cudaMemcpy(d_B, B, b_A, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        ScoreIn in{d_k1, d_k2, d_m, n, d_A, d_B, (float)(1.0 / (double)(sigma * sigma)), d_s, d_in};
        if (homography) k_ransac_score<true><<<K, kScoreThreads>>>(in);
        else k_ransac_score<false><<<K, kScoreThreads>>>(in);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(scores, d_s, b_s, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && inliers && b_in) e = cudaMemcpy(inliers, d_in, b_in, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return FBE_E_CUDA; }
    return FBE_OK;
}

}  // namespace fbe

extern "C" {

int fbe_check_homography(const fbe_keypoint* kps1, const fbe_keypoint* kps2, const int32_t* matches, int32_t n, const float* H21,
                         const float* H12, int32_t K, float sigma, int32_t device, float* scores, uint8_t* inliers) {
    return fbe::run_score(true, kps1, kps2, matches, n, H21, H12, K, sigma, device, scores, inliers);
}

int fbe_check_fundamental(const fbe_keypoint* kps1, const fbe_keypoint* kps2, const int32_t* matches, int32_t n, const float* F21,
                          int32_t K, float sigma, int32_t device, float* scores, uint8_t* inliers) {
    return fbe::run_score(false, kps1, kps2, matches, n, F21, nullptr, K, sigma, device, scores, inliers);
}

}  // extern "C"
