// Frame::isInFrustum (src/Frame.cc:435-491) for a whole local map at once: the producer of the query set
// of SearchByProjection(Frame&, vector<MapPoint*>&, th).  One thread per map point; everything is a handful of float /
// double operations whose order is the one OpenCV executes for the reference's cv::Mat expressions (pinned against
// cv2 4.13, tests/test_frustum.py):
//   Pc   = mRcw*P + mtcw    -> cv::gemm small-matrix path: t = a0*b0 + a1*b1 + a2*b2 in float, left to right, then
//                              (float)((double)t*1.0 + (double)c*1.0)
//   PO   = P - mOw          -> float subtraction
//   dist = cv::norm(PO)     -> sqrt of the double sum of squares, rounded to float on assignment
//   viewCos = PO.dot(Pn)/dist -> double dot product / (double)dist, rounded to float on assignment
//   MapPoint::PredictScale  -> ceil(logf(mfMaxDistance / dist) / mfLogScaleFactor) clamped to [0, mnScaleLevels)
// The library is built with -fmad=false, so none of the float expressions below contracts into an FMA.
#include <cmath>
#include "fbe_internal.cuh"

namespace fbe {

struct FrustumOut { uint8_t* in_view; float* proj; float* proj_xr; int32_t* level; float* view_cos; };

__global__ void k_is_in_frustum(const fbe_frustum_view v, const float* __restrict__ pos, const float* __restrict__ normal,
                                const float* __restrict__ min_dist, const float* __restrict__ max_dist, int n, float cos_limit,
                                FrustumOut o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float P0 = pos[3 * i], P1 = pos[3 * i + 1], P2 = pos[3 * i + 2];
    float Pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float t = v.Rcw[3 * r] * P0 + v.Rcw[3 * r + 1] * P1 + v.Rcw[3 * r + 2] * P2;
        Pc[r] = (float)((double)t + (double)v.tcw[r]);
    }
    bool ok = !(Pc[2] < 0.0f);                                        // :449-451
    const float invz = 1.0f / Pc[2];
    const float u = v.fx * Pc[0] * invz + v.cx;
    const float w = v.fy * Pc[1] * invz + v.cy;
    ok = ok && !(u < v.min_x || u > v.max_x) && !(w < v.min_y || w > v.max_y);
    const float maxD = 1.2f * max_dist[i], minD = 0.8f * min_dist[i];
    const float PO0 = P0 - v.Ow[0], PO1 = P1 - v.Ow[1], PO2 = P2 - v.Ow[2];
    const float dist = (float)sqrt((double)PO0 * (double)PO0 + (double)PO1 * (double)PO1 + (double)PO2 * (double)PO2);
    ok = ok && !(dist < minD || dist > maxD);
    const double dot = (double)PO0 * (double)normal[3 * i] + (double)PO1 * (double)normal[3 * i + 1] + (double)PO2 * (double)normal[3 * i + 2];
    const float view_cos = (float)(dot / (double)dist);
    ok = ok && !(view_cos < cos_limit);
    int lvl = 0;
    if (ok) {
        const float ratio = max_dist[i] / dist;
        lvl = (int)ceilf((float)log((double)ratio) / v.log_scale_factor);   // correctly rounded float log
        if (lvl < 0) lvl = 0;
        else if (lvl >= v.n_levels) lvl = v.n_levels - 1;
    }
    if (o.in_view) o.in_view[i] = ok ? 1 : 0;
    if (o.proj) { o.proj[2 * i] = ok ? u : 0.f; o.proj[2 * i + 1] = ok ? w : 0.f; }
    if (o.proj_xr) o.proj_xr[i] = ok ? u - v.mbf * invz : 0.f;
    if (o.level) o.level[i] = lvl;
    if (o.view_cos) o.view_cos[i] = ok ? view_cos : 0.f;
}

}  // namespace fbe

using namespace fbe;

extern "C" int fbe_is_in_frustum(const fbe_frustum_view* v, const float* pos, const float* normal, const float* min_dist,
                                 const float* max_dist, int32_t n, float viewing_cos_limit, int32_t device, uint8_t* in_view,
                                 float* proj, float* proj_xr, int32_t* level, float* view_cos) {
    if (!v || n < 0 || (n > 0 && (!pos || !normal || !min_dist || !max_dist)) || v->n_levels < 1) return FBE_E_INVALID;
    if (n == 0) return FBE_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device: this library has no CPU path"); return FBE_E_CUDA; }
    FBE_CUDA(cudaSetDevice(device));
    // one allocation: inputs (8 floats per point) then outputs (u8 + 5 x 4 bytes per point)
    const size_t N = (size_t)n, in_bytes = N * 8 * sizeof(float), out_bytes = N * 24;
    uint8_t* d = nullptr;
    FBE_CUDA(cudaMalloc(&d, in_bytes + out_bytes));
    float* d_pos = reinterpret_cast<float*>(d);
    float *d_nrm = d_pos + 3 * N, *d_min = d_nrm + 3 * N, *d_max = d_min + N;
    float* d_proj = reinterpret_cast<float*>(d + in_bytes);
    float *d_xr = d_proj + 2 * N, *d_cos = d_xr + N;
    int32_t* d_lvl = reinterpret_cast<int32_t*>(d_cos + N);
    uint8_t* d_in = reinterpret_cast<uint8_t*>(d_lvl + N);
    cudaError_t e = cudaMemcpy(d_pos, pos, 3 * N * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_nrm, normal, 3 * N * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_min, min_dist, N * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_max, max_dist, N * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        FrustumOut o{d_in, d_proj, d_xr, d_lvl, d_cos};
        k_is_in_frustum<<<(n + 127) / 128, 128>>>(*v, d_pos, d_nrm, d_min, d_max, n, viewing_cos_limit, o);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && in_view) e = cudaMemcpy(in_view, d_in, N, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && proj) e = cudaMemcpy(proj, d_proj, 2 * N * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && proj_xr) e = cudaMemcpy(proj_xr, d_xr, N * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && level) e = cudaMemcpy(level, d_lvl, N * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && view_cos) e = cudaMemcpy(view_cos, d_cos, N * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return FBE_E_CUDA; }
    return FBE_OK;
}
