// Phase 2 kernels: the daily convolution chain (CalcSol.py / cuda_lib.py).
//
// Semantics reproduced: the state lives on the reference's P x P torus,
// P = dom_len + max_shape//2 (CalcSol.py:20-21); one day is a circular
// convolution mod P with the wrap-shifted kernel (:58-66), the boundary flag is
// max(pad) > 1e-8 (:36-37) and a flagged state is truncated to the domain
// before the next day (:200-201).
//
// Method: P is data dependent and generally not smooth (1121 = 19*59,
// 4279 = 11*389), so instead of a length-P transform the step is evaluated as a
// LINEAR convolution on a 7-smooth torus N >= P + 2m (m = kernel radius) and
// folded mod P in real space -- identical to the circular convolution mod P in
// exact arithmetic, for any P.  One step is three streaming passes:
//
//   k_rows_fwd   real rows of the state  -> half-spectrum rows (two real rows
//                per complex FFT), written TRANSPOSED (Yt[kc][row])
//   k_cols       per spectral column: forward FFT, multiply with the kernel's
//                column spectrum, inverse FFT -- all three fused in shared
//                memory (Yt -> Wt), columns are contiguous in HBM
//   k_rows_inv   half-spectrum rows -> real rows, fold mod P, write the new
//                state and the per-row boundary/threshold statistics
//
// plus k_kernel_rows (row spectra of the day's kernel), k_step_finalize
// (flag / sum / count -> control block) and the emit kernels.
#pragma once
#include "fft_smem.cuh"

// launch-bound knobs (tools/chainbench.cu builds variants with -D)
#ifndef PKB_ROWS_T
#define PKB_ROWS_T 256
#endif
#ifndef PKB_ROWS_B
#define PKB_ROWS_B 1
#endif
#ifndef PKB_COLS_T
#define PKB_COLS_T 224
#endif
#ifndef PKB_COLS_B
#define PKB_COLS_B 1
#endif
#define PKB_ROWS_LB __launch_bounds__(PKB_ROWS_T, PKB_ROWS_B)
#define PKB_COLS_LB __launch_bounds__(PKB_COLS_T, PKB_COLS_B)
#define PKB_COLS_TMAX PKB_COLS_T

namespace pkb {

struct ChainDims {
    int D;      // domain side (dom_len)
    int P;      // reference torus side (pad_shape)
    int N;      // FFT torus side, 7-smooth, >= P + 2*mmax
    int Nc;     // N/2 + 1 spectral columns kept
    int ldS;    // leading dimension (doubles) of real states
    int ldY;    // leading dimension (complex) of Yt  (>= P)
    int ldW;    // leading dimension (complex) of Wt  (>= N)
    int ldK;    // leading dimension (complex) of kernel row spectra (>= 2*mmax+1)
};

struct ChainCtrl {
    int trunc;   // state is zero outside [0,D)^2: only that block is read
    int flag;    // boundary flag of the last step
    int pad_[2];
};

struct StepMeta {   // one per emitted solution
    double padmax, ksum, add, vmin;
    long long kcnt;
    int flag;
    int pad_;
};

// ---------------------------------------------------------------------------
// Hermitian split of the transform of z = a + i b (a, b real rows) at bin k
__device__ __forceinline__ void unpack_pair(const cplx* x, const FftPlan& plan, int N, int k, cplx& A, cplx& B) {
    const cplx zk = x[swz(__ldg(&plan.perm[k]))];
    const cplx zn = x[swz(__ldg(&plan.perm[k == 0 ? 0 : N - k]))];
    A = cmake(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
    B = cmake(0.5 * (zk.y + zn.y), 0.5 * (zn.x - zk.x));
}

// grid = ceil(P/2), block = T, dyn smem = Npad complex
__global__ void PKB_ROWS_LB k_rows_fwd(const double* __restrict__ S, ChainDims d, const ChainCtrl* __restrict__ ctrl,
                                                 cplx* __restrict__ Yt, FftPlan plan) {
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    const int lim = ctrl->trunc ? d.D : d.P;
    const int r0 = 2 * blockIdx.x;
    if (r0 >= lim) return;
    const int r1 = (r0 + 1 < lim) ? r0 + 1 : -1;
    const int tid = threadIdx.x, T = blockDim.x;
    const double* s0 = S + (size_t)r0 * d.ldS;
    const double* s1 = S + (size_t)(r1 >= 0 ? r1 : r0) * d.ldS;
    auto ld = [&](int j) -> cplx {
        if (j >= lim) return cmake(0.0, 0.0);
        return cmake(s0[j], r1 >= 0 ? s1[j] : 0.0);
    };
    fft_dif_from(x, plan, tid, T, ld);
    const int N = d.N;
    for (int k = tid; k < d.Nc; k += T) {
        cplx A, B;
        unpack_pair(x, plan, N, k, A, B);
        cplx* dst = Yt + (size_t)k * d.ldY + r0;
        dst[0] = A;
        if (r1 >= 0) dst[1] = B;
    }
}

// Row spectra of the wrap-shifted kernel (CalcSol.py:58-64 on the N torus).
// K: dense (Wk x Wk) window centred on the release cell, support radius m.
// Krt[kc][q], q = dy for dy in [0,m], q = dy + 2m+1 for dy in [-m,-1].
// grid = m+1 (row pairs), block = T, dyn smem = Npad complex
__global__ void __launch_bounds__(256) k_kernel_rows(const double* __restrict__ K, int Wk, int m, ChainDims d, cplx* __restrict__ Krt,
                                                    FftPlan plan) {
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    const int nq = 2 * m + 1;
    const int q0 = 2 * blockIdx.x;
    if (q0 >= nq) return;
    const int q1 = (q0 + 1 < nq) ? q0 + 1 : -1;
    const int tid = threadIdx.x, T = blockDim.x;
    const int ck = Wk / 2;
    const int dy0 = q0 <= m ? q0 : q0 - nq;
    const int dy1 = q1 < 0 ? 0 : (q1 <= m ? q1 : q1 - nq);
    const double* k0 = K + (size_t)(ck + dy0) * Wk + ck;
    const double* k1 = K + (size_t)(ck + dy1) * Wk + ck;
    const int N = d.N;
    auto ld = [&](int j) -> cplx {
        int dx;
        if (j <= m) dx = j;
        else if (j >= N - m) dx = j - N;
        else return cmake(0.0, 0.0);
        return cmake(k0[dx], q1 >= 0 ? k1[dx] : 0.0);
    };
    fft_dif_from(x, plan, tid, T, ld);
    for (int k = tid; k < d.Nc; k += T) {
        cplx A, B;
        unpack_pair(x, plan, N, k, A, B);
        cplx* dst = Krt + (size_t)k * d.ldK;
        dst[q0] = A;
        if (q1 >= 0) dst[q1] = B;
    }
}

// ---------------------------------------------------------------------------
// Last forward stage of a column, fused with the spectral multiply and the
// first inverse stage.  Thread t owns the final-stage blocks t, t + T, ... (KB of
// them, R_last points each): the filter's spectrum of those blocks stays in
// registers (K) while the same shared-memory buffer is reused for the state
// column, so the filter spectrum and the product never touch shared memory.
#define PKB_COLS_KMAX 32   // KB * R_last <= 32 complex registers

template <int RL, int KB>
__device__ __forceinline__ void cols_final_filter(const cplx* x, int tid, int T, int nbl, cplx (&K)[PKB_COLS_KMAX]) {
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
        const int j = tid + kb * T;
        if (j < nbl) {
            cplx v[RL];
#pragma unroll
            for (int q = 0; q < RL; ++q) v[q] = x[swz(j * RL + q)];
            dft<RL>(v);
#pragma unroll
            for (int q = 0; q < RL; ++q) K[kb * RL + q] = v[q];
        }
    }
}
template <int RL, int KB>
__device__ __forceinline__ void cols_final_state(cplx* x, int tid, int T, int nbl, const cplx (&K)[PKB_COLS_KMAX]) {
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
        const int j = tid + kb * T;
        if (j < nbl) {
            cplx v[RL];
#pragma unroll
            for (int q = 0; q < RL; ++q) v[q] = x[swz(j * RL + q)];
            dft<RL>(v);
#pragma unroll
            for (int q = 0; q < RL; ++q) v[q] = cmul_f(v[q], K[kb * RL + q]);
            idft<RL>(v);
#pragma unroll
            for (int q = 0; q < RL; ++q) x[swz(j * RL + q)] = v[q];
        }
    }
}

// the (R_last, KB) pairs the planner may pick (pkb200.cu: plan_cols)
#define PKB_COLS_SWITCH(RL, KB, CALL)                                                                     \
    switch ((RL) * 8 + (KB)) {                                                                            \
        case 2 * 8 + 1: { CALL(2, 1); } break;  case 2 * 8 + 2: { CALL(2, 2); } break;                    \
        case 2 * 8 + 3: { CALL(2, 3); } break;  case 2 * 8 + 4: { CALL(2, 4); } break;                    \
        case 3 * 8 + 1: { CALL(3, 1); } break;  case 3 * 8 + 2: { CALL(3, 2); } break;                    \
        case 3 * 8 + 3: { CALL(3, 3); } break;  case 3 * 8 + 4: { CALL(3, 4); } break;                    \
        case 4 * 8 + 1: { CALL(4, 1); } break;  case 4 * 8 + 2: { CALL(4, 2); } break;                    \
        case 4 * 8 + 3: { CALL(4, 3); } break;  case 4 * 8 + 4: { CALL(4, 4); } break;                    \
        case 5 * 8 + 1: { CALL(5, 1); } break;  case 5 * 8 + 2: { CALL(5, 2); } break;                    \
        case 5 * 8 + 3: { CALL(5, 3); } break;  case 5 * 8 + 4: { CALL(5, 4); } break;                    \
        case 7 * 8 + 1: { CALL(7, 1); } break;  case 7 * 8 + 2: { CALL(7, 2); } break;                    \
        case 7 * 8 + 3: { CALL(7, 3); } break;  case 7 * 8 + 4: { CALL(7, 4); } break;                    \
        case 8 * 8 + 1: { CALL(8, 1); } break;  case 8 * 8 + 2: { CALL(8, 2); } break;                    \
        case 8 * 8 + 3: { CALL(8, 3); } break;  case 8 * 8 + 4: { CALL(8, 4); } break;                    \
        case 9 * 8 + 1: { CALL(9, 1); } break;  case 9 * 8 + 2: { CALL(9, 2); } break;                    \
        default: { CALL(9, 3); } break;                                                                   \
    }

// grid = Nc, block = plan.cols_threads, dyn smem = Npad complex, 2 CTAs/SM.
// Per spectral column: forward FFT of the filter column (inputs straight from
// Krt, only 2m+1 of them non-zero), forward FFT of the state column (inputs
// straight from Yt), product, inverse FFT, rows needed by the fold straight to Wt.
__global__ void PKB_COLS_LB k_cols(const cplx* __restrict__ Yt, const cplx* __restrict__ Krt, int m, ChainDims d,
                                                const ChainCtrl* __restrict__ ctrl, cplx* __restrict__ Wt, FftPlan plan) {
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    const int c = blockIdx.x;
    const int lim = ctrl->trunc ? d.D : d.P;
    const int tid = threadIdx.x, T = blockDim.x;
    const int N = d.N, nq = 2 * m + 1;
    const cplx* ycol = Yt + (size_t)c * d.ldY;
    const cplx* kcol = Krt + (size_t)c * d.ldK;
    cplx* wcol = Wt + (size_t)c * d.ldW;
    const int L = plan.nstage, RL = plan_radix(plan, L - 1), nbl = N / RL, KB = plan.cols_kb;
    const int hi = d.P + m;   // rows [0, P+m) and [N-m, N) are needed by the fold
    auto ld_filter = [&](int i) -> cplx {
        if (i <= m) return kcol[i];
        if (i >= N - m) return kcol[i - (N - nq)];
        return cmake(0.0, 0.0);
    };
    auto ld_state = [&](int i) -> cplx { return i < lim ? ycol[i] : cmake(0.0, 0.0); };
    auto st_out = [&](int i, cplx v) {
        if (i < hi || i >= N - m) wcol[i] = v;
    };
    cplx K[PKB_COLS_KMAX];
    const int R0 = plan_radix(plan, 0);

    // phase 0: filter column spectrum -> K registers; phase 1: state column
    // forward, product, inverse (one copy of the stage code serves both)
    for (int phase = 0; phase < 2; ++phase) {
        auto ld = [&](int i) -> cplx { return phase ? ld_state(i) : ld_filter(i); };
        if (L == 1) {
            for (int i = tid; i < N; i += T) x[swz(i)] = ld(i);
        } else {
            fft_stage_first_dispatch(x, R0, N, plan.tw, tid, T, ld);
            int M = N / R0, off = 0;
            for (int s = 1; s < L - 1; ++s) {
                __syncthreads();
                const int R = plan_radix(plan, s);
                plan_stage_fwd(x, plan, R, M, off, tid, T);
                off += stage_tw_size(R, M);
                M /= R;
            }
        }
        __syncthreads();
        if (phase == 0) {
#define PKB_CALL_(RR, KK) cols_final_filter<RR, KK>(x, tid, T, nbl, K)
            PKB_COLS_SWITCH(RL, KB, PKB_CALL_)
#undef PKB_CALL_
        } else {
#define PKB_CALL_(RR, KK) cols_final_state<RR, KK>(x, tid, T, nbl, K)
            PKB_COLS_SWITCH(RL, KB, PKB_CALL_)
#undef PKB_CALL_
        }
        __syncthreads();
    }
    if (L == 1) {
        for (int i = tid; i < N; i += T) st_out(i, x[swz(i)]);
    } else {
        int M = RL;
        int off = plan_tw_offset(plan, L - 2);
        for (int s = L - 2; s >= 1; --s) {
            const int R = plan_radix(plan, s);
            M *= R;
            off -= stage_tw_size(R, M);
            plan_stage_inv(x, plan, R, M, off, tid, T);
            __syncthreads();
        }
        fft_stage_last_inv_dispatch(x, R0, N, plan.tw, tid, T, st_out);
    }
}

// ---------------------------------------------------------------------------
// Per-row statistics written by k_rows_inv, reduced by k_step_finalize.
struct RowStats {
    double padmax, ksum, vmin;
    int kcnt;
    int pad_;
};

__device__ __forceinline__ double fold_col(const cplx* x, int c, int P, int N, int m, bool imag) {
    cplx a = x[swz(c)];
    double v = imag ? a.y : a.x;
    if (c < m) { cplx b = x[swz(c + P)]; v += imag ? b.y : b.x; }
    if (c >= P - m) { cplx b = x[swz(c - P + N)]; v += imag ? b.y : b.x; }
    return v;
}

// grid = 2m + ceil((P-2m)/2), block = T, dyn smem = Npad complex
__global__ void PKB_ROWS_LB k_rows_inv(const cplx* __restrict__ Wt, int m, ChainDims d, double* __restrict__ Sout, RowStats* __restrict__ rstat,
                           double negval, FftPlan plan) {
    PKB_DYN_SMEM(raw);
    PKB_SHARED(double, red, 1024);
    cplx* x = reinterpret_cast<cplx*>(raw);
    const int P = d.P, N = d.N, D = d.D;
    const int job = blockIdx.x;
    int ra, rb, out_a, out_b;
    bool fold;
    if (job < m) { fold = true; out_a = job; out_b = -1; ra = job; rb = job + P; }
    else if (job < 2 * m) { const int t = job - m; fold = true; out_a = P - m + t; out_b = -1; ra = out_a; rb = N - m + t; }
    else {
        const int r = m + 2 * (job - 2 * m);
        fold = false; out_a = r; ra = r;
        out_b = (r + 1 < P - m) ? r + 1 : -1; rb = out_b;
    }
    const int tid = threadIdx.x, T = blockDim.x;
    const cplx zero = cmake(0.0, 0.0);
    // zero the padding slots, then scatter the Hermitian pair Z = A + iB
    for (int i = N + tid; i < plan.Npad; i += T) x[swz(i)] = zero;
    for (int k = tid; k < d.Nc; k += T) {
        const cplx a = Wt[(size_t)k * d.ldW + ra];
        const cplx b = rb >= 0 ? Wt[(size_t)k * d.ldW + rb] : zero;
        const int nk = N - k;
        if (k == 0 || nk == k) {
            x[swz(__ldg(&plan.perm[k]))] = cmake(a.x, b.x);        // self-conjugate bins are real
        } else {
            x[swz(__ldg(&plan.perm[k]))] = cmake(a.x - b.y, a.y + b.x);    // A + iB
            x[swz(__ldg(&plan.perm[nk]))] = cmake(a.x + b.y, b.x - a.y);   // conj(A) + i conj(B)
        }
    }
    __syncthreads();
    fft_dit_inv(x, plan, tid, T);
    const double scale = 1.0 / ((double)N * (double)N);
    const int nout = (fold || out_b < 0) ? 1 : 2;
    for (int o = 0; o < nout; ++o) {
        const int r = o ? out_b : out_a;
        double* dst = Sout + (size_t)r * d.ldS;
        double pmax = -INFINITY, ks = 0.0, vmn = INFINITY;
        int kc = 0;
        for (int c = tid; c < P; c += T) {
            double v;
            if (fold) v = fold_col(x, c, P, N, m, false) + fold_col(x, c, P, N, m, true);
            else v = fold_col(x, c, P, N, m, o == 1);
            v *= scale;
            dst[c] = v;
            if (r >= D || c >= D) pmax = fmax(pmax, v);
            else {
                vmn = fmin(vmn, v);
                if (!(v < negval)) { ks += v; kc += 1; }
            }
        }
        const double bpmax = block_max(pmax, red);
        const double bks = block_sum(ks, red);
        const double bkc = block_sum((double)kc, red);
        const double bmn = block_min(vmn, red);
        if (tid == 0) {
            RowStats rs;
            rs.padmax = bpmax; rs.ksum = bks; rs.vmin = bmn; rs.kcnt = (int)bkc; rs.pad_ = 0;
            rstat[r] = rs;
        }
    }
}

// grid = 1, block = 256.  Fixed-order tree over the P rows.
__global__ void k_step_finalize(const RowStats* __restrict__ rstat, ChainDims d, ChainCtrl* __restrict__ ctrl, StepMeta* __restrict__ meta,
                                int apply_trunc) {
    PKB_SHARED(double, red, 256);
    const int tid = threadIdx.x, T = blockDim.x;
    double pmax = -INFINITY, ks = 0.0, kc = 0.0, vmn = INFINITY;
    // contiguous chunks per thread so the summation order over rows is fixed
    const int chunk = (d.P + T - 1) / T;
    for (int r = tid * chunk; r < d.P && r < (tid + 1) * chunk; ++r) {
        const RowStats rs = rstat[r];
        pmax = fmax(pmax, rs.padmax);
        if (r < d.D) { ks += rs.ksum; kc += (double)rs.kcnt; vmn = fmin(vmn, rs.vmin); }
    }
    const double bp = block_max(pmax, red);
    const double bs = block_sum(ks, red);
    const double bc = block_sum(kc, red);
    const double bm = block_min(vmn, red);
    if (tid == 0) {
        const int flag = bp > 1e-8 ? 1 : 0;          // CalcSol.py:36-37
        meta->padmax = bp; meta->ksum = bs; meta->kcnt = (long long)bc; meta->vmin = bm;
        meta->add = (1.0 - bs) / bc;                 // CalcSol.py:135
        meta->flag = flag;
        ctrl->flag = flag;
        // a fresh convolution result is a full P x P state; it becomes a
        // truncated one only where the caller applies CalcSol.py:200-201
        ctrl->trunc = apply_trunc ? flag : 0;
    }
}

// get_cursol's "Re-fft" decision taken after the fact (cuda_lib.py:130-136)
__global__ void k_apply_trunc(ChainCtrl* __restrict__ ctrl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) ctrl->trunc = ctrl->flag;
}
__global__ void k_set_ctrl(ChainCtrl* __restrict__ ctrl, int trunc, int flag) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { ctrl->trunc = trunc; ctrl->flag = flag; }
}

// out[r][c] = r_small_vals(S[:D,:D], prob_model) densely (CalcSol.py:112-136)
// grid = D, block = 256
// strict != 0: keep v > negval (cuda_lib.py:117-119) instead of !(v < negval)
__global__ void k_emit_dense(const double* __restrict__ S, ChainDims d, const StepMeta* __restrict__ meta, double negval, int prob_model,
                             int strict, double* __restrict__ out) {
    const int r = blockIdx.x;
    const double add = prob_model ? meta->add : 0.0;
    const double* src = S + (size_t)r * d.ldS;
    double* dst = out + (size_t)r * d.D;
    for (int c = threadIdx.x; c < d.D; c += blockDim.x) {
        const double v = src[c];
        const bool keep = strict ? (v > negval) : (v != 0.0 && !(v < negval));
        dst[c] = keep ? v + add : 0.0;
    }
}

// Zero everything outside [0,D)^2 of a flagged state (the zero padding of the
// reference's re-FFT, CalcSol.py:200-201 / cuda_lib.py:132-135).  grid = P
__global__ void k_zero_pad(double* __restrict__ S, ChainDims d, const ChainCtrl* __restrict__ ctrl) {
    if (!ctrl->trunc) return;
    const int r = blockIdx.x;
    double* row = S + (size_t)r * d.ldS;
    const int c0 = r < d.D ? d.D : 0;
    for (int c = c0 + threadIdx.x; c < d.P; c += blockDim.x) row[c] = 0.0;
}

// plain copy of the domain block (un-thresholded), grid = D
__global__ void k_copy_domain(const double* __restrict__ S, ChainDims d, double* __restrict__ out) {
    const int r = blockIdx.x;
    for (int c = threadIdx.x; c < d.D; c += blockDim.x) out[(size_t)r * d.D + c] = S[(size_t)r * d.ldS + c];
}

// Place a dense centred kernel window (Wk x Wk, radius used: m) into a zeroed
// state at the domain centre (Run.py:454-458).  grid = 2m+1, block = 128
__global__ void k_place_kernel(const double* __restrict__ K, int Wk, int m, ChainDims d, double* __restrict__ S) {
    const int ck = Wk / 2, cd = d.D / 2;
    const int dy = (int)blockIdx.x - m;
    for (int t = threadIdx.x; t < 2 * m + 1; t += blockDim.x) {
        const int dx = t - m;
        S[(size_t)(cd + dy) * d.ldS + cd + dx] = K[(size_t)(ck + dy) * Wk + ck + dx];
    }
}

// Load a dense D x D host-provided state block into a zeroed P x P state. grid = D
__global__ void k_load_state(const double* __restrict__ A, ChainDims d, double* __restrict__ S) {
    const int r = blockIdx.x;
    for (int c = threadIdx.x; c < d.D; c += blockDim.x) S[(size_t)r * d.ldS + c] = A[(size_t)r * d.D + c];
}

// Population model output (CalcSol.py:271-274,303-306,322-323):
//   tot = (sum_d S_d * w_d) * r_number ; r_small_vals(tot) ; centre += extra
// grid = D, block = 256
struct CohortArgs {
    const double* S[16];
    double w[16];
    int n;
};
// first_day != 0: CalcSol.py:236-237, r_small_vals(r_spread[0]) * r_number * dist(1)
// (threshold on the probability, product order (v * r_number) * w).
// pre (optional): un-thresholded weighted sum (parity export).
__global__ void k_emit_population(CohortArgs ca, ChainDims d, double r_number, double centre_extra, int add_centre, double negval,
                                  int first_day, double* __restrict__ out, double* __restrict__ pre) {
    const int r = blockIdx.x;
    const int mid = d.D / 2;
    for (int c = threadIdx.x; c < d.D; c += blockDim.x) {
        double v;
        if (first_day) {
            const double p = ca.S[0][(size_t)r * d.ldS + c];
            const double t = (p != 0.0 && !(p < negval)) ? p : 0.0;
            v = (t * r_number) * ca.w[0];
            if (pre) pre[(size_t)r * d.D + c] = v;
        } else {
            double acc = 0.0;
            for (int k = 0; k < ca.n; ++k) acc += ca.S[k][(size_t)r * d.ldS + c] * ca.w[k];
            v = acc * r_number;
            if (pre) pre[(size_t)r * d.D + c] = v;
            v = (v != 0.0 && !(v < negval)) ? v : 0.0;
        }
        if (add_centre && r == mid && c == mid) v += centre_extra;
        out[(size_t)r * d.D + c] = v;
    }
}

// out[day][k] = G[day][cells[k]]   grid = ndays, block = 256
__global__ void k_sample(const double* __restrict__ G, int D, const int* __restrict__ cells, int K, double* __restrict__ out) {
    const double* g = G + (size_t)blockIdx.x * D * D;
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        out[(size_t)blockIdx.x * K + k] = g[(size_t)cells[2 * k] * D + cells[2 * k + 1]];
}

// Single-vector transform through the shared-memory FFT (diagnostics).
// grid = 1, block = T, dyn smem = Npad complex
__global__ void __launch_bounds__(256) k_fft_test(const cplx* __restrict__ in, cplx* __restrict__ out, int inverse, FftPlan plan) {
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    const int tid = threadIdx.x, T = blockDim.x, N = plan.N;
    for (int i = tid; i < plan.Npad; i += T) x[swz(i)] = cmake(0.0, 0.0);
    __syncthreads();
    if (!inverse) {
        for (int i = tid; i < N; i += T) x[swz(i)] = in[i];
        __syncthreads();
        fft_dif(x, plan, tid, T);
        for (int k = tid; k < N; k += T) out[k] = x[swz(__ldg(&plan.perm[k]))];
    } else {
        for (int k = tid; k < N; k += T) x[swz(__ldg(&plan.perm[k]))] = in[k];
        __syncthreads();
        fft_dit_inv(x, plan, tid, T);
        for (int i = tid; i < N; i += T) out[i] = x[swz(i)];
    }
}

// Row-major compaction of dense [ndays][D][D] grids into COO (scipy.sparse.coo_matrix
// ordering).  Pass 1: per-row non-zero counts; pass 2 (after an exclusive scan
// over rows): ordered write.  grid = ndays*D, block = 256
__global__ void k_row_nnz(const double* __restrict__ G, int D, int* __restrict__ rownnz) {
    PKB_SHARED(double, red, 256);
    int n = 0;
    for (int c = threadIdx.x; c < D; c += blockDim.x) n += (G[(size_t)blockIdx.x * D + c] != 0.0) ? 1 : 0;
    const double t = block_sum((double)n, red);
    if (threadIdx.x == 0) rownnz[blockIdx.x] = (int)t;
}
// Exclusive scan of the per-row counts of all days: rowoff[day*D + r] is the
// global COO position of row r of that day; dayoff[day] the start of the day,
// dayoff[ndays] the grand total.  grid = 1, block = 1024; each thread owns a
// contiguous chunk of the ndays*D rows.
__global__ void __launch_bounds__(1024) k_row_scan(const int* __restrict__ rownnz, int D, int ndays, long long* __restrict__ rowoff, long long* __restrict__ dayoff) {
    PKB_SHARED(long long, part, 1024);
    const long long n = (long long)D * ndays;
    const int tid = threadIdx.x, T = blockDim.x;
    const long long chunk = (n + T - 1) / T;
    const long long lo = tid * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
    long long acc = 0;
    for (long long i = lo; i < hi; ++i) acc += rownnz[i];
    part[tid] = acc;
    __syncthreads();
    if (tid == 0) {
        long long run = 0;
        for (int t = 0; t < T; ++t) { const long long v = part[t]; part[t] = run; run += v; }
        dayoff[ndays] = run;
    }
    __syncthreads();
    acc = part[tid];
    for (long long i = lo; i < hi; ++i) {
        rowoff[i] = acc;
        if (i % D == 0) dayoff[i / D] = acc;
        acc += rownnz[i];
    }
}
// grid = ndays*D, block = 256
__global__ void k_coo_write(const double* __restrict__ G, int D, const long long* __restrict__ rowoff, int* __restrict__ rows,
                            int* __restrict__ cols, double* __restrict__ vals) {
    PKB_SHARED(int, cnt, 256);
    PKB_SHARED(int, basepos, 1);
    const int r = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) basepos[0] = 0;
    __syncthreads();
    for (int c0 = 0; c0 < D; c0 += T) {
        const int c = c0 + tid;
        const double v = c < D ? G[(size_t)r * D + c] : 0.0;
        const int nz = v != 0.0 ? 1 : 0;
        cnt[tid] = nz;
        __syncthreads();
        // inclusive scan (Hillis-Steele) over the block
        for (int s = 1; s < T; s <<= 1) {
            int add = tid >= s ? cnt[tid - s] : 0;
            __syncthreads();
            cnt[tid] += add;
            __syncthreads();
        }
        if (nz) {
            const long long pos = rowoff[r] + basepos[0] + cnt[tid] - 1;
            rows[pos] = r % D; cols[pos] = c; vals[pos] = v;
        }
        __syncthreads();
        if (tid == 0) basepos[0] += cnt[T - 1];
        __syncthreads();
    }
}

// Direct circular convolution mod P for small kernels (stencil path, K8).
// out[r][c] = sum_{dy,dx} K(dy,dx) * S[(r-dy) mod P][(c-dx) mod P], state read
// with the same truncation rule as k_rows_fwd.  grid = (ceil(P/32), ceil(P/8)),
// block = (32, 8), dyn smem = (8+2m)*(32+2m) + (2m+1)^2 doubles.
__global__ void k_stencil(const double* __restrict__ S, const double* __restrict__ K, int Wk, int m, ChainDims d,
                          const ChainCtrl* __restrict__ ctrl, double* __restrict__ Sout) {
    PKB_DYN_SMEM(raw);
    double* tile = reinterpret_cast<double*>(raw);
    const int TW = 32 + 2 * m, TH = 8 + 2 * m, nk = 2 * m + 1;
    double* ker = tile + TW * TH;
    const int P = d.P, lim = ctrl->trunc ? d.D : d.P;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 8;
    const int ck = Wk / 2;
    for (int i = tid; i < nk * nk; i += 256) {
        const int dy = i / nk - m, dx = i % nk - m;
        ker[i] = K[(size_t)(ck + dy) * Wk + ck + dx];
    }
    for (int i = tid; i < TW * TH; i += 256) {
        const int ty = i / TW, tx = i - ty * TW;
        int r = r0 + ty - m, c = c0 + tx - m;
        r %= P; if (r < 0) r += P;
        c %= P; if (c < 0) c += P;
        tile[i] = (r < lim && c < lim) ? S[(size_t)r * d.ldS + c] : 0.0;
    }
    __syncthreads();
    const int r = r0 + threadIdx.y, c = c0 + threadIdx.x;
    if (r < P && c < P) {
        double acc = 0.0;
        for (int dy = -m; dy <= m; ++dy)
            for (int dx = -m; dx <= m; ++dx)
                acc = fma(ker[(dy + m) * nk + dx + m], tile[(threadIdx.y + m - dy) * TW + threadIdx.x + m - dx], acc);
        Sout[(size_t)r * d.ldS + c] = acc;
    }
}

// Row statistics of a state (used after the stencil path). grid = P, block = 256
__global__ void k_row_stats(const double* __restrict__ S, ChainDims d, RowStats* __restrict__ rstat, double negval) {
    PKB_SHARED(double, red, 256);
    const int r = blockIdx.x;
    double pmax = -INFINITY, ks = 0.0, vmn = INFINITY;
    int kc = 0;
    for (int c = threadIdx.x; c < d.P; c += blockDim.x) {
        const double v = S[(size_t)r * d.ldS + c];
        if (r >= d.D || c >= d.D) pmax = fmax(pmax, v);
        else {
            vmn = fmin(vmn, v);
            if (!(v < negval)) { ks += v; kc += 1; }
        }
    }
    const double bpmax = block_max(pmax, red);
    const double bks = block_sum(ks, red);
    const double bkc = block_sum((double)kc, red);
    const double bmn = block_min(vmn, red);
    if (threadIdx.x == 0) {
        RowStats rs;
        rs.padmax = bpmax; rs.ksum = bks; rs.vmin = bmn; rs.kcnt = (int)bkc; rs.pad_ = 0;
        rstat[r] = rs;
    }
}

}  // namespace pkb
