// Platform layer: real CUDA by default.  -DPKB_EMUL (tests/emul only) swaps in
// the CPU fiber emulation so kernel logic can be unit-tested without a GPU; the
// product library is never built that way (see tests/emul/emul_cuda.h).
#pragma once
#ifdef PKB_EMUL
#include "emul_cuda.h"
#define PKB_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch_kernel(dim3(grid), dim3(block), (size_t)(smem), kern, __VA_ARGS__)
#define PKB_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#define PKB_SHARED(type, name, n) static thread_local type name[n]
#define PKB_IS_EMUL 1
#else
#include <cuda_runtime.h>
#define PKB_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<dim3(grid), dim3(block), (size_t)(smem), (stream)>>>(__VA_ARGS__)
#define PKB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define PKB_SHARED(type, name, n) __shared__ type name[n]
#define PKB_IS_EMUL 0
#endif

#include <cmath>
#include <cstdint>

namespace pkb {

typedef double2 cplx;

__host__ __device__ __forceinline__ cplx cmake(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__host__ __device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
    return cmake(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.x, -a.y); }

// Two adjacent complex128 values as ONE 256-bit global access (sm_100 LDG/STG.256):
// the transposed spectra are written / read as 32-byte row pairs, one L1 wavefront
// per pair instead of two.  p must be 32-byte aligned.
__device__ __forceinline__ void st_pair(cplx* p, cplx a, cplx b) {
#if PKB_IS_EMUL
    p[0] = a;
    p[1] = b;
#else
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
#endif
}
// hint: bring the 128-byte line at p into L2 (persistent kernels prefetch their NEXT job's inputs)
__device__ __forceinline__ void prefetch_l2(const void* p) {
#if !PKB_IS_EMUL
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
__device__ __forceinline__ void ld_pair(const cplx* p, cplx& a, cplx& b) {
#if PKB_IS_EMUL
    a = p[0];
    b = p[1];
#else
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
#endif
}

// Block-wide reductions through shared memory (scratch >= blockDim.x doubles).
// Fixed tree => deterministic results for a given launch shape.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int tid = threadIdx.x, n = blockDim.x;
    __syncthreads();
    scratch[tid] = v;
    __syncthreads();
    for (int s = 1; s < n; s <<= 1) {
        int i = 2 * s * tid;
        if (i + s < n) scratch[i] += scratch[i + s];
        __syncthreads();
    }
    double r = scratch[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
    const int tid = threadIdx.x, n = blockDim.x;
    __syncthreads();
    scratch[tid] = v;
    __syncthreads();
    for (int s = 1; s < n; s <<= 1) {
        int i = 2 * s * tid;
        if (i + s < n) scratch[i] = fmax(scratch[i], scratch[i + s]);
        __syncthreads();
    }
    double r = scratch[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_min(double v, double* scratch) {
    return -block_max(-v, scratch);
}

}  // namespace pkb
