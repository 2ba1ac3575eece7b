// Device-only timing harness for one chain step (tuning tool, not part of the product).
//
// Builds the library sources into one executable so that kernel variants can be
// selected with -D macros and timed back to back on the GPU box:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo [-D...] -o tools/chainbench_X tools/chainbench.cu
//   tools/chainbench_X [dom_len] [filter_side] [steps] [fft_threads]
// Prints the per-kernel mean device time (CUDA events around every launch) and a
// checksum of the state so that variants can be compared for agreement.
#include "../parasitoids_b200/csrc/pkb200.cu"

#include <cstdlib>
#include <random>

// pure shared-memory FFT rate: every CTA transforms its own buffer forward and back `reps` times
__global__ void PKB_ROWS_LB k_fft_core(cplx* out, int reps, FftPlan plan) {
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    cplx* tws = x + plan.N;
    const int tid = threadIdx.x, T = blockDim.x;
    fft_load_twiddles(tws, plan, tid, T);
    for (int i = tid; i < plan.N; i += T) x[i] = cmake(1.0 / (i + 1), 0.5 / (i + 2));
    __syncthreads();
    const double sc = 1.0 / plan.N;
    for (int r = 0; r < reps; ++r) {
        fft_forward_from(x, tws, plan, tid, T, SmemLoad{x});
        fft_inverse_to(x, tws, plan, tid, T, [&](int i, cplx v) { x[i] = cmake(v.x * sc, v.y * sc); });
        __syncthreads();
    }
    if (tid == 0) out[blockIdx.x] = x[blockIdx.x % plan.N];
}

static int core_bench(pkb_ctx* ctx, int N, int T, int reps) {
    FftPlan plan;
    if (get_plan(ctx, N, &plan)) { fprintf(stderr, "%s\n", pkb_last_error()); return 1; }
    opt_in_smem(k_fft_core);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fft_core, T, fft_smem_bytes(plan));
    const int grid = occ * ctx->sm_count;
    cplx* out = nullptr;
    cudaMalloc(&out, sizeof(cplx) * grid);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_fft_core<<<grid, T, fft_smem_bytes(plan), ctx->stream>>>(out, 2, plan);
    cudaEventRecord(a, ctx->stream);
    k_fft_core<<<grid, T, fft_smem_bytes(plan), ctx->stream>>>(out, reps, plan);
    cudaEventRecord(b, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    printf("core N=%d radices", N);
    for (int s = 0; s < plan.nstage; ++s) printf(" %d", plan_radix(plan, s));
    const double ntr = 2.0 * reps * grid;   // forward + inverse
    printf("  T=%d occ=%d: %.2f us per transform per SM (%.0f cycles at 1.965 GHz), err %s\n", T, occ,
           1000.0 * ms / (ntr / ctx->sm_count), 1965.0 * 1000.0 * ms / (ntr / ctx->sm_count), cudaGetErrorString(cudaGetLastError()));
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "core")) {
        pkb_ctx* c0 = nullptr;
        if (pkb_create(0, &c0)) { fprintf(stderr, "%s\n", pkb_last_error()); return 1; }
        return core_bench(c0, argc > 2 ? atoi(argv[2]) : 4704, argc > 3 ? atoi(argv[3]) : 256, argc > 4 ? atoi(argv[4]) : 20);
    }
    const int D = argc > 1 ? atoi(argv[1]) : 4097;
    const int k = argc > 2 ? atoi(argv[2]) : 361;
    const int steps = argc > 3 ? atoi(argv[3]) : 10;
    const int T = argc > 4 ? atoi(argv[4]) : 0;
    pkb_ctx* ctx = nullptr;
    if (pkb_create(0, &ctx)) { fprintf(stderr, "%s\n", pkb_last_error()); return 1; }
    if (T && pkb_set_option(ctx, "fft_threads", T)) { fprintf(stderr, "%s\n", pkb_last_error()); return 1; }
    pkb_chain* ch = nullptr;
    if (pkb_chain_create(ctx, D, k, &ch)) { fprintf(stderr, "%s\n", pkb_last_error()); return 1; }
    int P = 0, N = 0;
    pkb_chain_dims(ch, nullptr, &P, &N);
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::vector<double> A((size_t)D * D), B((size_t)k * k);
    double sa = 0, sb = 0;
    for (auto& v : A) { v = U(rng); sa += v; }
    for (auto& v : B) { v = U(rng); sb += v; }
    for (auto& v : A) v /= sa;
    for (auto& v : B) v /= sb;
    if (pkb_chain_set_state(ch, A.data())) { fprintf(stderr, "%s\n", pkb_last_error()); return 1; }
    for (int i = 0; i < 3; ++i)
        if (pkb_chain_conv(ch, B.data(), k)) { fprintf(stderr, "%s\n", pkb_last_error()); return 1; }
    pkb_profile_reset(ctx);
    pkb_profile_enable(ctx, 1);
    for (int i = 0; i < steps; ++i)
        if (pkb_chain_conv(ch, B.data(), k)) { fprintf(stderr, "%s\n", pkb_last_error()); return 1; }
    pkb_profile_enable(ctx, 0);
    const char* names[] = {"k_kernel_rows", "k_rows_fwd", "k_cols", "k_rows_inv", "k_step_finalize"};
    double tot = 0;
    printf("D=%d k=%d P=%d N=%d steps=%d:", D, k, P, N, steps);
    for (const char* nm : names) {
        long long c = 0;
        double ms = 0;
        pkb_profile_get(ctx, nm, &c, &ms);
        if (c) { printf("  %s %.1f us", nm + 2, 1000.0 * ms / c); tot += 1000.0 * ms / c; }
    }
    std::vector<double> S((size_t)P * P);
    pkb_chain_get_state(ch, S.data());
    double sum = 0, sq = 0;
    for (double v : S) { sum += v; sq += v * v; }
    printf("  | total %.1f us | sum %.15e sumsq %.15e\n", tot, sum, sq);
    pkb_chain_destroy(ch);
    pkb_destroy(ctx);
    return 0;
}
