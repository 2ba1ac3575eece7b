// cuFFT cross-check / yardstick for the chain's hand-written FFT (tuning tool, NOT part of the product;
// BASELINE.json north_star: "cuFFT used only as a cross-check").
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/cufft_yardstick tools/cufft_yardstick.cu -lcufft
//   tools/cufft_yardstick [N ...]          (default: 4704 4480 4277 1458 1121)
//
// For every side N it times, with CUDA events after warm-up, one chain step done the cuFFT way on an N x N real
// fp64 grid: D2Z of the kernel, pointwise product with the resident state spectrum, Z2D of the product -- the
// reference's own per-day sequence (CalcSol.py:58-66 fftconv2 + :28-41 ifft2; its state stays spectral) -- and,
// separately, D2Z + Z2D alone.  Prints one JSON line per N.
#include <cuda_runtime.h>
#include <cufft.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void k_mul(cufftDoubleComplex* a, const cufftDoubleComplex* b, size_t n, double scale) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double ar = a[i].x, ai = a[i].y, br = b[i].x, bi = b[i].y;
        a[i].x = (ar * br - ai * bi) * scale;
        a[i].y = (ar * bi + ai * br) * scale;
    }
}

#define CK(x) do { if ((x) != 0) { fprintf(stderr, "%s failed (%d) line %d\n", #x, (int)(x), __LINE__); return 1; } } while (0)

static int run(int N, int reps) {
    const size_t nr = (size_t)N * N, nc = (size_t)N * (N / 2 + 1);
    double *state, *kern, *out;
    cufftDoubleComplex *shat, *khat, *phat;
    CK(cudaMalloc(&state, nr * 8)); CK(cudaMalloc(&kern, nr * 8)); CK(cudaMalloc(&out, nr * 8));
    CK(cudaMalloc(&shat, nc * 16)); CK(cudaMalloc(&khat, nc * 16)); CK(cudaMalloc(&phat, nc * 16));
    std::vector<double> h(nr);
    for (size_t i = 0; i < nr; ++i) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 / nr;
    CK(cudaMemcpy(state, h.data(), nr * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(kern, h.data(), nr * 8, cudaMemcpyHostToDevice));
    cufftHandle pf, pi;
    size_t wf = 0, wi = 0;
    CK(cufftCreate(&pf)); CK(cufftCreate(&pi));
    CK(cufftMakePlan2d(pf, N, N, CUFFT_D2Z, &wf));
    CK(cufftMakePlan2d(pi, N, N, CUFFT_Z2D, &wi));
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    CK(cufftExecD2Z(pf, state, shat));
    for (int w = 0; w < 3; ++w) { CK(cufftExecD2Z(pf, kern, khat)); CK(cufftExecZ2D(pi, khat, out)); }
    CK(cudaDeviceSynchronize());
    // (a) transforms alone
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) { CK(cufftExecD2Z(pf, kern, khat)); CK(cufftExecZ2D(pi, khat, out)); }
    cudaEventRecord(e1);
    // (b) the reference's chain step: kernel D2Z, state_hat *= kernel_hat, Z2D of a copy (Z2D destroys its input)
    for (int r = 0; r < reps; ++r) {
        CK(cufftExecD2Z(pf, kern, khat));
        k_mul<<<148 * 8, 256>>>(shat, khat, nc, 1.0);
        CK(cudaMemcpyAsync(phat, shat, nc * 16, cudaMemcpyDeviceToDevice));
        CK(cufftExecZ2D(pi, phat, out));
    }
    cudaEventRecord(e2);
    CK(cudaDeviceSynchronize());
    float t_a = 0, t_b = 0;
    cudaEventElapsedTime(&t_a, e0, e1);
    cudaEventElapsedTime(&t_b, e1, e2);
    printf("{\"bench\": \"cufft_fp64_2d\", \"N\": %d, \"d2z_plus_z2d_us\": %.1f, \"chain_step_us\": %.1f, \"workspace_mb\": %.1f, \"reps\": %d}\n", N,
           1000.0 * t_a / reps, 1000.0 * t_b / reps, (wf + wi) / 1e6, reps);
    cufftDestroy(pf); cufftDestroy(pi);
    cudaFree(state); cudaFree(kern); cudaFree(out); cudaFree(shat); cudaFree(khat); cudaFree(phat);
    return 0;
}

int main(int argc, char** argv) {
    std::vector<int> sizes;
    for (int i = 1; i < argc; ++i) sizes.push_back(atoi(argv[i]));
    if (sizes.empty()) sizes = {4704, 4480, 4277, 1458, 1121};
    for (int N : sizes)
        if (run(N, N > 2000 ? 20 : 100)) return 1;
    return 0;
}
