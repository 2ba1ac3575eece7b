// CPU emulation of the small CUDA subset used by parasitoids_b200/csrc.
//
// TEST INFRASTRUCTURE ONLY.  This header lets the *same* kernel sources be
// compiled with g++ (-DPKB_EMUL) into tests/emul/libpkb200_emul.so so that
// indexing / control-flow logic can be checked against the oracle in the
// GPU-less build container before GPU minutes are spent.  The product library
// (parasitoids_b200/libpkb200.so) never includes this file, and nothing under
// parasitoids_b200/ can load the emulation library: there is no CPU fallback.
//
// Model: every CUDA block runs on one OS thread; its CUDA threads are ucontext
// fibers executed round-robin; __syncthreads() yields to the block scheduler,
// which resumes the fibers only after every live fiber has arrived.
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <tuple>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __align__(n) alignas(n)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline int4 make_int4(int a, int b, int c, int d) { int4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

namespace emu {
struct Fiber {
    ucontext_t ctx;
    dim3 tid;
    bool done;
};
struct BlockCtx {
    dim3 bid, bdim, gdim;
    unsigned char* dyn;
    Fiber* cur;
    ucontext_t sched;
    const std::function<void()>* body;
};
extern thread_local BlockCtx* tl_block;
void sync();
void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F f) {
    std::function<void()> body(f);
    launch_impl(grid, block, smem, body);
}
// kernel<<<grid, block, smem>>>(args...): arguments are evaluated and converted
// to the kernel's parameter types at the launch site, like a real launch
template <class... KArgs, class... Args>
inline void launch_kernel(dim3 grid, dim3 block, size_t smem, void (*kern)(KArgs...), Args... args) {
    std::tuple<KArgs...> pack(static_cast<KArgs>(args)...);
    launch(grid, block, smem, [=]() { std::apply(kern, pack); });
}
inline unsigned char* dyn_smem() { return tl_block->dyn; }
}  // namespace emu

#define threadIdx (emu::tl_block->cur->tid)
#define blockIdx (emu::tl_block->bid)
#define blockDim (emu::tl_block->bdim)
#define gridDim (emu::tl_block->gdim)
#define __syncthreads() emu::sync()

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline double atomicAdd(double* addr, double val) {
    uint64_t* p = reinterpret_cast<uint64_t*>(addr);
    uint64_t old = __atomic_load_n(p, __ATOMIC_RELAXED), nw;
    double o;
    do {
        memcpy(&o, &old, 8);
        double n = o + val;
        memcpy(&nw, &n, 8);
    } while (!__atomic_compare_exchange_n(p, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return o;
}
static inline int atomicAdd(int* addr, int val) { return __atomic_fetch_add(addr, val, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* addr, unsigned long long val) {
    return __atomic_fetch_add(addr, val, __ATOMIC_RELAXED);
}
static inline long long __double2ll_rn(double v) { return llrint(v); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int atomicOr(int* addr, int val) { return __atomic_fetch_or(addr, val, __ATOMIC_RELAXED); }
static inline int atomicMax(int* addr, int val) {
    int old = __atomic_load_n(addr, __ATOMIC_RELAXED);
    while (old < val && !__atomic_compare_exchange_n(addr, &old, val, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
// warp shuffle for block-uniform call sites (every thread of the block calls it the
// same number of times): exchange through a per-block buffer between two fiber barriers
static inline double __shfl_down_sync(unsigned, double v, int delta) {
    static thread_local double ex[1024];
    const unsigned t = threadIdx.x;
    ex[t] = v;
    emu::sync();
    const unsigned lane = t & 31u;
    const double r = (lane + (unsigned)delta < 32u && t + (unsigned)delta < blockDim.x) ? ex[t + delta] : v;
    emu::sync();
    return r;
}
static inline int2 make_int2(int a, int b) { int2 r; r.x = a; r.y = b; return r; }
static inline void sincospi(double x, double* s, double* c) {
    *s = sin(M_PI * x);
    *c = cos(M_PI * x);
}

// ---- runtime API shims -------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize, cudaFuncAttributePreferredSharedMemoryCarveout };
static inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emulated failure" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = nullptr;
    if (posix_memalign(p, 256, n ? n : 256)) return cudaErrorMemoryAllocation;
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                            cudaStream_t = nullptr) {
    for (size_t r = 0; r < height; ++r) memcpy((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
enum { cudaStreamNonBlocking = 1 };
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
#define cudaErrorNotReady ((cudaError_t)600)
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 3; return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
struct cudaFuncAttributes { size_t sharedSizeBytes; };
template <class F> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { a->sharedSizeBytes = 0; return cudaSuccess; }
