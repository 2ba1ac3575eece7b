// Device microbenchmarks that set the non-HBM ceilings quoted in DESIGN.md:
//   dfma   fp64 FMA issue rate (the bound of the FFT butterflies and of the BVN lattice)
//   exp    fp64 exp() rate (phase 1: one exp per Gauss-Legendre node)
//   smem   shared-memory bandwidth with 16-byte (complex128) accesses
//   copy   HBM copy bandwidth (cross-check of MEASURED_PEAKS.json)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, int iters) {
    double a[4];
    for (int i = 0; i < 4; ++i) a[i] = -1.0 - threadIdx.x * 1e-3 - i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = -1.0 - exp(a[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3];
}

__global__ void k_smem(double2* out, int iters) {
    extern __shared__ double2 sm[];
    const int n = blockDim.x * 4;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2 v = sm[(threadIdx.x + j * blockDim.x + it) % n];
            acc.x += v.x; acc.y += v.y;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

template <class F>
static float timeit(F f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* d;
    cudaMalloc(&d, sizeof(double) * 2 * sms * 8 * 1024);
    {
        const int iters = 4096, blocks = sms * 8, threads = 512;
        float ms = timeit([&] { k_dfma<<<blocks, threads>>>(d, iters); }, 5);
        double fl = 2.0 * 8 * iters * (double)blocks * threads;
        printf("{\"bench\": \"dfma\", \"tflops\": %.2f, \"fma_per_clk_per_sm\": %.1f, \"sms\": %d}\n", fl / ms / 1e9,
               fl / 2 / (ms * 1e-3) / sms / (p.clockRate * 1e3), sms);
    }
    {
        const int iters = 512, blocks = sms * 8, threads = 512;
        float ms = timeit([&] { k_exp<<<blocks, threads>>>(d, iters); }, 5);
        double n = 4.0 * iters * (double)blocks * threads;
        printf("{\"bench\": \"exp_f64\", \"gexp_per_s\": %.1f}\n", n / ms / 1e6);
    }
    {
        const int iters = 4096, blocks = sms * 4, threads = 512;
        float ms = timeit([&] { k_smem<<<blocks, threads, threads * 4 * sizeof(double2)>>>((double2*)d, iters); }, 5);
        double bytes = 16.0 * 4 * iters * (double)blocks * threads;
        printf("{\"bench\": \"smem_ld128\", \"tbytes_per_s\": %.2f, \"bytes_per_clk_per_sm\": %.1f}\n", bytes / ms / 1e9,
               bytes / (ms * 1e-3) / sms / (p.clockRate * 1e3));
    }
    {
        const size_t n = (size_t)1 << 27;   // 2 GiB each way
        double2 *a, *b;
        cudaMalloc(&a, n * sizeof(double2));
        cudaMalloc(&b, n * sizeof(double2));
        cudaMemset(a, 0, n * sizeof(double2));
        float ms = timeit([&] { k_copy<<<sms * 16, 512>>>(a, b, n); }, 5);
        printf("{\"bench\": \"copy\", \"gbytes_per_s\": %.1f}\n", 2.0 * n * sizeof(double2) / ms / 1e6);
    }
    return 0;
}
