// SASS probe: which byte-SIMD intrinsics are single instructions on sm_100a (cuobjdump -sass)
__global__ void k_max(const unsigned* a, const unsigned* b, unsigned* o) { int i = threadIdx.x; o[i] = __vmaxu4(a[i], b[i]); }
__global__ void k_min(const unsigned* a, const unsigned* b, unsigned* o) { int i = threadIdx.x; o[i] = __vminu4(a[i], b[i]); }
__global__ void k_cmpgt(const unsigned* a, const unsigned* b, unsigned* o) { int i = threadIdx.x; o[i] = __vcmpgtu4(a[i], b[i]); }
__global__ void k_subus(const unsigned* a, const unsigned* b, unsigned* o) { int i = threadIdx.x; o[i] = __vsubus4(a[i], b[i]); }
__global__ void k_absdiff(const unsigned* a, const unsigned* b, unsigned* o) { int i = threadIdx.x; o[i] = __vabsdiffu4(a[i], b[i]); }
__global__ void k_addus(const unsigned* a, const unsigned* b, unsigned* o) { int i = threadIdx.x; o[i] = __vaddus4(a[i], b[i]); }
__global__ void k_setgt(const unsigned* a, const unsigned* b, unsigned* o) { int i = threadIdx.x; o[i] = __vsetgtu4(a[i], b[i]); }
