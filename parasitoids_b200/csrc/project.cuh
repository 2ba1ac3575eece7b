// Likelihood projection on the device (Bayes_funcs.py:20-180): what the Bayesian drivers read from a forward
// solve -- expected emergence per collection / observation date at the release-field grid points and in the
// sentinel fields (popdensity_to_emergence), and the population at the grid points on the observation days
// (popdensity_grid).  All of it is a small ordered linear map of the model at K sample cells:
//
//   S[day][set]   = sum of the model over the cells of a set (one grid cell, or the cells of a sentinel field)
//   G[row][group] = ((0 + S[d1][set] w1) + S[d2][set] w2) + ...     one emergence day: `emerg_proj[n, e] += ...`
//                                                                   accumulated over oviposition days (:58-71)
//   out[row]      = sum over the row's groups                      `emerg_proj[:, a:b].sum(axis=1)` (:82-85)
//
// The two sums written "sum" are numpy reductions over contiguous data, i.e. numpy's pairwise summation; it is
// restated below so that the device result is the reference's to the last bit, not just to rounding.
#pragma once
#include "pkb_platform.cuh"

namespace pkb {

struct ProjTables {
    const int* set_ptr;      // [nsets + 1]  -> set_cells
    const int* set_cells;    // indices into the K sample cells
    const int* row_ptr;      // [nrows + 1]  -> groups
    const int* grp_ptr;      // [ngroups + 1] -> terms
    const int* term_day;
    const int* term_set;
    const double* term_w;
    int nrows;
};

// numpy's pairwise sum of n values get(0..n-1) (numpy/core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum):
// < 8 values sequentially; up to 128 with eight running sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and the
// remainder added one by one; beyond that split at n/2 rounded down to a multiple of 8, recursively.
template <class F>
__device__ double np_pairwise(F get, int lo, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += get(lo + i);
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = get(lo + j);
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += get(lo + i + j);
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += get(lo + i);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise(get, lo, n2) + np_pairwise(get, lo + n2, n - n2);
}

// samples: [nprop][nd][K]; out: [nprop][nrows].  One thread per (proposal, row).
__global__ void k_project(ProjTables t, const double* __restrict__ samples, int nd, int K, int nprop, double* __restrict__ out) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)nprop * t.nrows) return;
    const int p = (int)(id / t.nrows), row = (int)(id - (long long)p * t.nrows);
    const double* smp = samples + (size_t)p * nd * K;
    auto set_sum = [&](int day, int set) -> double {
        const int c0 = t.set_ptr[set], nc = t.set_ptr[set + 1] - c0;
        const double* m = smp + (size_t)day * K;
        if (nc == 1) return m[t.set_cells[c0]];
        return np_pairwise([&](int i) { return m[t.set_cells[i]]; }, c0, nc);
    };
    auto group = [&](int g) -> double {
        double acc = 0.0;
        for (int k = t.grp_ptr[g]; k < t.grp_ptr[g + 1]; ++k) acc += set_sum(t.term_day[k], t.term_set[k]) * t.term_w[k];
        return acc;
    };
    const int g0 = t.row_ptr[row], ng = t.row_ptr[row + 1] - g0;
    out[id] = np_pairwise(group, g0, ng);
}

}  // namespace pkb
