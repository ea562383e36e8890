// POPC-pipe peak on this GPU: register-resident xor + popc + add chains, 8 independent accumulators per thread.
// Prints G POPC/s and, from that, the 256-bit Hamming pairs/s ceiling (8 POPC per pair).
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(unsigned* out, int iters, unsigned seed) {
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    for (int i = 0; i < iters; ++i) {
        s0 += __popc(a0 ^ s7); s1 += __popc(a1 ^ s0); s2 += __popc(a2 ^ s1); s3 += __popc(a3 ^ s2);
        s4 += __popc(a4 ^ s3); s5 += __popc(a5 ^ s4); s6 += __popc(a6 ^ s5); s7 += __popc(a7 ^ s6);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
}
__global__ void k_indep(unsigned* out, int iters, unsigned seed) {
    unsigned a[8], s[8];
    for (int j = 0; j < 8; ++j) { a[j] = seed * (2 * j + 3) + threadIdx.x; s[j] = 0; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += __popc(a[j]); a[j] = a[j] * 1664525u + 1013904223u; }
    }
    unsigned t = 0;
    for (int j = 0; j < 8; ++j) t += s[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
int main() {
    unsigned* d; cudaMalloc(&d, 148 * 16 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int v = 0; v < 2; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (v == 0) k<<<148 * 16, 256>>>(d, iters, 12345u); else k_indep<<<148 * 16, 256>>>(d, iters, 12345u);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double popc = 148.0 * 16 * 256 * (double)iters * 8;
        printf("variant %d: %.3f ms, %.1f G POPC/s (%.2f per clk per SM at 1.965 GHz) -> %.1f G Hamming-256 pairs/s ceiling\n", v, ms,
               popc / ms / 1e6, popc / ms / 1e6 / 148 / 1.965, popc / ms / 1e6 / 8);
    }
    return 0;
}
