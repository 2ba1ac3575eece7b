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

int main(int argc, char** argv) {
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
