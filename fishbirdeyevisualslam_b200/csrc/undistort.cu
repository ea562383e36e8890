// Frame::UndistortKeyPoints (src/Frame.cc:638-669) and Frame::ComputeImageBounds (:741-795): the step between extraction
// and grid assignment.  The reference calls cv::fisheye::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(),
// mK) -- third-party arithmetic, pinned here to OpenCV 4.13 (tests/golden/undistort.npz, generated with cv2):
//   pw = (p - c) / f;  theta_d = clamp(|pw|, -pi/2, pi/2);  Newton on theta*(1 + k1 t^2 + k2 t^4 + k3 t^6 + k4 t^8) = theta_d
//   (<= 10 iterations, stop at |fix| < 1e-8);  scale = tan(theta) / theta_d;  out = f * (pw * scale) + c,
//   (-1e6, -1e6) when the iteration did not converge or theta flipped sign.  All in double, result rounded to float.
// One thread per point; fp64 throughput is irrelevant at a few thousand points per frame.  The keypoint record is copied
// and only pt is replaced, like the reference does (kp = mvKeys[i]; kp.pt = ...).
#include <cmath>
#include "fbe_internal.cuh"

namespace fbe {

__host__ __device__ inline void fisheye_undistort_point(double px, double py, double fx, double fy, double cx, double cy,
                                                        double k0, double k1, double k2, double k3, float* ox, float* oy) {
    const double wx = (px - cx) / fx, wy = (py - cy) / fy;
    double theta_d = sqrt(wx * wx + wy * wy);
    const double half_pi = 3.1415926535897932384626433832795 / 2.;
    theta_d = fmin(fmax(-half_pi, theta_d), half_pi);
    bool converged = false;
    double theta = theta_d, scale = 0.0;
    if (fabs(theta_d) > 1e-8) {
        for (int j = 0; j < 10; ++j) {
            const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
            const double a = k0 * t2, b = k1 * t4, c = k2 * t6, d = k3 * t8;
            const double fix = (theta * (1 + a + b + c + d) - theta_d) / (1 + 3 * a + 5 * b + 7 * c + 9 * d);
            theta = theta - fix;
            if (fabs(fix) < 1e-8) { converged = true; break; }
        }
        scale = tan(theta) / theta_d;
    } else {
        converged = true;
    }
    const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
    if (converged && !flipped) {
        const double ux = wx * scale, uy = wy * scale;
        // RR = P = K (as double); pr = K * (ux, uy, 1); fi = pr.xy / pr.z with pr.z == 1
        *ox = (float)((fx * ux + 0.0 * uy + cx * 1.0) / (0.0 * ux + 0.0 * uy + 1.0));
        *oy = (float)((0.0 * ux + fy * uy + cy * 1.0) / (0.0 * ux + 0.0 * uy + 1.0));
    } else {
        *ox = -1000000.0f; *oy = -1000000.0f;
    }
}

__global__ void k_undistort(const fbe_keypoint* __restrict__ in, int n, float fx, float fy, float cx, float cy, float k0, float k1,
                            float k2, float k3, fbe_keypoint* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fbe_keypoint kp = in[i];
    float x, y;
    fisheye_undistort_point((double)kp.x, (double)kp.y, (double)fx, (double)fy, (double)cx, (double)cy, (double)k0, (double)k1,
                            (double)k2, (double)k3, &x, &y);
    kp.x = x; kp.y = y;
    out[i] = kp;
}

// the same per-point arithmetic over the device-resident keypoints of a batch (n_arr[b] valid entries per image)
__global__ void k_undistort_batch(const fbe_keypoint* __restrict__ in, const int* __restrict__ n_arr, int stride, float fx, float fy,
                                  float cx, float cy, float k0, float k1, float k2, float k3, fbe_keypoint* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_arr[b]) return;
    fbe_keypoint kp = in[(size_t)b * stride + i];
    float x, y;
    fisheye_undistort_point((double)kp.x, (double)kp.y, (double)fx, (double)fy, (double)cx, (double)cy, (double)k0, (double)k1,
                            (double)k2, (double)k3, &x, &y);
    kp.x = x; kp.y = y;
    out[(size_t)b * stride + i] = kp;
}

int launch_undistort_batch(const fbe_keypoint* in, const int* n_arr, int stride, int nimg, const float K[4], const float D[4],
                           fbe_keypoint* out, cudaStream_t st) {
    dim3 grid((stride + 127) / 128, nimg);
    FBE_CUDA(launch_dep(k_undistort_batch, grid, dim3(128), 0, st, in, n_arr, stride, K[0], K[1], K[2], K[3], D[0], D[1], D[2], D[3], out));
    count_launch();
    FBE_CUDA(cudaGetLastError());
    return FBE_OK;
}

}  // namespace fbe

using namespace fbe;

extern "C" {

int fbe_undistort_keypoints(const fbe_keypoint* kps, int32_t n, const float K[4], const float D[4], int32_t device, fbe_keypoint* out) {
    if (n < 0 || !K || !D || (n > 0 && (!kps || !out))) return FBE_E_INVALID;
    if (n == 0) return FBE_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device: this library has no CPU path"); return FBE_E_CUDA; }
    FBE_CUDA(cudaSetDevice(device));
    if (D[0] == 0.0f) {                       // mvKeysUn = mvKeys (:640-644)
        if (out != kps) for (int i = 0; i < n; ++i) out[i] = kps[i];
        return FBE_OK;
    }
    fbe_keypoint *d_in = nullptr, *d_out = nullptr;
    FBE_CUDA(cudaMalloc(&d_in, (size_t)n * sizeof(fbe_keypoint)));
    if (cudaMalloc(&d_out, (size_t)n * sizeof(fbe_keypoint)) != cudaSuccess) { cudaFree(d_in); set_error("cudaMalloc failed"); return FBE_E_CUDA; }
    cudaError_t e = cudaMemcpy(d_in, kps, (size_t)n * sizeof(fbe_keypoint), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_undistort<<<(n + 127) / 128, 128>>>(d_in, n, K[0], K[1], K[2], K[3], D[0], D[1], D[2], D[3], d_out);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)n * sizeof(fbe_keypoint), cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return FBE_E_CUDA; }
    return FBE_OK;
}

int fbe_image_bounds(int32_t cols, int32_t rows, const float K[4], const float D[4], int32_t device, float bounds[4]) {
    if (!K || !D || !bounds || cols <= 0 || rows <= 0) return FBE_E_INVALID;
    if (D[0] == 0.0f) {                       // :789-794
        bounds[0] = 0.0f; bounds[1] = (float)cols; bounds[2] = 0.0f; bounds[3] = (float)rows;
        return FBE_OK;
    }
    fbe_keypoint c[4] = {}, u[4];
    c[0].x = 0.f; c[0].y = 0.f; c[1].x = (float)cols; c[1].y = 0.f; c[2].x = 0.f; c[2].y = (float)rows; c[3].x = (float)cols; c[3].y = (float)rows;
    const int rc = fbe_undistort_keypoints(c, 4, K, D, device, u);
    if (rc != FBE_OK) return rc;
    // :762-779 -- note the reference initialises the maxima with numeric_limits<float>::min() (the smallest POSITIVE float)
    float mnx = 3.402823466e+38f, mxx = 1.175494351e-38f, mny = 3.402823466e+38f, mxy = 1.175494351e-38f;
    for (int i = 0; i < 4; ++i) {
        if (u[i].x < mnx) mnx = u[i].x;
        if (u[i].x > mxx) mxx = u[i].x;
        if (u[i].y < mny) mny = u[i].y;
        if (u[i].y > mxy) mxy = u[i].y;
    }
    bounds[0] = mnx; bounds[1] = mxx; bounds[2] = mny; bounds[3] = mxy;
    return FBE_OK;
}

}  // extern "C"
