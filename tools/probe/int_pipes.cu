// Integer instruction throughput on this GPU, register-resident chains with 8 independent accumulators per thread:
// IMAD, IADD3, LOP3, PRMT, VIMNMX3.U16x2 (DPX), IDP4A (__dp4a), IDP2A (__dp2a_lo), and two 50/50 mixes that tell
// whether a pair of instructions shares a pipe (mix rate == single rate) or not (mix rate ~ 2x).
// Prints results per clock per SM (clock taken from the device's current SM clock).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>

// nvcc fuses the two HMNMX2 of a nested __hmin2 into ONE 3-input VHMNMX on sm_100a
__device__ __forceinline__ unsigned hmin3(unsigned a, unsigned b, unsigned c) {
    const __half2 r = __hmin2(__hmin2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b)), *reinterpret_cast<const __half2*>(&c));
    return *reinterpret_cast<const unsigned*>(&r);
}

template <int OP>
__global__ void k(unsigned* out, int iters, unsigned seed) {
    unsigned a[8], s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = seed * (2 * j + 3) + threadIdx.x; s[j] = j; }
    const unsigned c = seed | 0x01020304u;
    const unsigned ch = (seed & 0x03FF03FFu) | 0x64006400u;                         // two positive normal fp16 values
    if (OP >= 18) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] = (a[j] & 0x03FF03FFu) | 0x64006400u; s[j] = 0x65FF65FFu - j; }
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) s[j] = s[j] * c + a[j];                                    // IMAD
            if (OP == 1) s[j] = s[j] + a[j] + c;                                    // IADD3
            if (OP == 2) s[j] = (s[j] & a[j]) ^ c;                                  // LOP3
            if (OP == 3) s[j] = __byte_perm(s[j], a[j], 0x5241);                    // PRMT
            if (OP == 4) s[j] = __vimin3_u16x2(s[j], a[j], c);                      // VIMNMX3.U16x2
            if (OP == 5) s[j] = __dp4a(a[j], c, s[j]);                              // IDP.4A
            if (OP == 6) s[j] = __dp2a_lo(a[j], c, s[j]);                           // IDP.2A
            if (OP == 7) { if (j & 1) s[j] = s[j] * c + a[j]; else s[j] = __vimin3_u16x2(s[j], a[j], c); }   // IMAD + VIMNMX3
            if (OP == 8) { if (j & 1) s[j] = __dp4a(a[j], c, s[j]); else s[j] = __vimin3_u16x2(s[j], a[j], c); }   // IDP4A + VIMNMX3
            if (OP == 9) { if (j & 1) s[j] = __dp4a(a[j], c, s[j]); else s[j] = s[j] * c + a[j]; }         // IDP4A + IMAD
            if (OP == 10) s[j] = __vimax3_s16x2(s[j], a[j], c);                     // VIMNMX3.S16x2
            if (OP == 11) s[j] = __funnelshift_r(s[j], a[j], 8);                    // SHF
            if (OP == 12) { __half2 x = *reinterpret_cast<__half2*>(&s[j]), y = *reinterpret_cast<__half2*>(&a[j]); x = __hmax2(x, y); s[j] = *reinterpret_cast<unsigned*>(&x) + 1u; }   // HMNMX2 (+1 keeps the chain live)
            if (OP == 13) s[j] = __umulhi(s[j], c) + a[j];                           // IMAD.HI
            if (OP == 14) s[j] = __float_as_uint(fmaxf(__uint_as_float(s[j]), __uint_as_float(a[j]))) ^ c;   // FMNMX
            if (OP == 15) s[j] = __vmaxs2(s[j], a[j]) + c;                      // VIMNMX.S16x2 (2-input)
            if (OP == 16) { if (j & 1) { __half2 x = *reinterpret_cast<__half2*>(&s[j]), y = *reinterpret_cast<__half2*>(&a[j]); x = __hmax2(x, y); s[j] = *reinterpret_cast<unsigned*>(&x); } else s[j] = __vimin3_u16x2(s[j], a[j], c); }   // HMNMX2 + VIMNMX3 mix
            if (OP == 17) s[j] = __popc(s[j] ^ a[j]) + c;                            // POPC
            if (OP == 18) s[j] = hmin3(s[j], a[j], ch);                               // VHMNMX (3-input half2 min, sm_100)
            if (OP == 19) { if (j & 1) s[j] = hmin3(s[j], a[j], ch); else s[j] = __vimin3_s16x2(s[j], a[j], ch); }   // VHMNMX + VIMNMX3
            if (OP == 20) { if (j & 1) s[j] = hmin3(s[j], a[j], ch); else s[j] = s[j] * c + a[j]; }                  // VHMNMX + IMAD
            if (OP == 21) { if (j % 3 == 0) s[j] = hmin3(s[j], a[j], ch); else s[j] = __vimin3_s16x2(s[j], a[j], ch); }   // 1 VHMNMX : 2 VIMNMX3
        }
    }
    unsigned t = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += s[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int OP>
static void run(const char* name, unsigned* d, double ghz) {
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<OP><<<148 * 8, 256>>>(d, iters, 12345u + rep);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double ops = 148.0 * 8 * 256 * (double)iters * 8;
    printf("%-22s %8.3f ms  %7.1f thread-ops/clk/SM  (%.2f warp-instr/clk/SM)\n", name, ms, ops / (ms * 1e-3) / 148 / (ghz * 1e9),
           ops / (ms * 1e-3) / 148 / (ghz * 1e9) / 32);
}

int main() {
    unsigned* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    printf("SM clock (max) %.3f GHz\n", ghz);
    run<0>("IMAD", d, ghz); run<1>("IADD3", d, ghz); run<2>("LOP3", d, ghz); run<3>("PRMT", d, ghz); run<11>("SHF", d, ghz);
    run<4>("VIMNMX3.U16x2", d, ghz); run<10>("VIMNMX3.S16x2", d, ghz); run<5>("IDP4A", d, ghz); run<6>("IDP2A", d, ghz);
    run<12>("HMNMX2(+IADD)", d, ghz); run<13>("IMAD.HI", d, ghz); run<14>("FMNMX(+LOP)", d, ghz); run<15>("VIMNMX.S16x2(+IADD)", d, ghz); run<16>("HMNMX2+VIMNMX3 mix", d, ghz); run<17>("POPC(+LOP+IADD)", d, ghz);
    run<18>("VHMNMX", d, ghz); run<19>("VHMNMX+VIMNMX3 mix", d, ghz); run<20>("VHMNMX+IMAD mix", d, ghz); run<21>("VHMNMX+2 VIMNMX3 mix", d, ghz);
    run<7>("IMAD+VIMNMX3 mix", d, ghz); run<8>("IDP4A+VIMNMX3 mix", d, ghz); run<9>("IDP4A+IMAD mix", d, ghz);
    return 0;
}
