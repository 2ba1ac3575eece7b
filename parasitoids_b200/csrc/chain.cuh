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
// exact arithmetic, for any P.  One step is three streaming passes, each a
// PERSISTENT kernel (grid = resident CTAs per SM x SM count, every CTA loops over
// its jobs; one job = one transform held in one shared-memory buffer):
//
//   k_rows_fwd   real rows of the state  -> half-spectrum rows (two real rows
//                per complex FFT), written into the tiled transposed array Yt
//   k_cols       per spectral column: forward FFT of the kernel column (parked
//                in an L2-resident scratch), forward FFT of the state column,
//                product, inverse FFT -- one CTA, one pass over the column
//                (Yt -> Wt)
//   k_rows_inv   half-spectrum rows -> real rows, fold mod P, write the new
//                state and the per-row boundary/threshold statistics; the CTA
//                that finishes last reduces them to the step's flag / sums
//
// plus k_kernel_rows(_batch) (row spectra of the daily kernels), the emission
// kernels (r_small_vals, cohort superposition), the COO compaction kernels and
// the direct stencil for tiny kernels.  While the state's exact support still
// fits inside the domain a step runs on a torus sized for that support window
// (ChainDims::win).  Whole-torus steps of the fused solve carry up to three
// refinements, all picked per step and all the same circular convolution mod P
// to rounding: the step's own torus N >= P + 2m for THAT day's m
// (pkb200.cu:step_torus); a second geometry N_t >= D + 2m that the kernels switch
// to on the device when the source state is truncated (TruncGeom); and, between
// two steps on one torus, the next step's forward row pass run by k_rows_inv on
// the row pair it still holds (Yt_next, ChainCtrl::fused).  DESIGN.md sections
// 3-5 give the layout and the numbers.
#pragma once
#include "fft_smem.cuh"

// launch bounds: at most 256 threads and 128 registers (2 x 256, 3 x 160 or 4 x 128 resident threads per SM,
// chosen per plan in pkb200.cu:get_plan); tools/chainbench.cu builds variants with -D
#ifndef PKB_ROWS_T
#define PKB_ROWS_T 256
#endif
#ifndef PKB_ROWS_B
#define PKB_ROWS_B 2
#endif
#ifndef PKB_COLS_T
#define PKB_COLS_T 256
#endif
#ifndef PKB_COLS_B
#define PKB_COLS_B 2
#endif
#ifdef PKB_MAXNREG      // explicit register cap instead of launch bounds (tuning builds: 3 x 224 threads need <= 96)
#define PKB_ROWS_LB __maxnreg__(PKB_MAXNREG)
#define PKB_COLS_LB __maxnreg__(PKB_MAXNREG)
#else
#define PKB_ROWS_LB __launch_bounds__(PKB_ROWS_T, PKB_ROWS_B)
#define PKB_COLS_LB __launch_bounds__(PKB_COLS_T, PKB_COLS_B)
#endif
#define PKB_COLS_TMAX PKB_COLS_T

namespace pkb {

struct ChainDims {
    int D;      // domain side (dom_len)
    int P;      // reference torus side (pad_shape)
    int N;      // FFT torus side, 7-smooth, >= P + 2*mmax
    int Nc;     // N/2 + 1 spectral columns kept
    int ldS;    // leading dimension (doubles) of real states
    int ldY;    // row slots per column tile of Yt  (>= P)
    int ldW;    // row slots per column tile of Wt  (>= N)
    int ldK;    // row slots per column tile of the kernel row spectra (>= 2*mmax+1)
    // Window mode (win != 0): the state is known to be EXACTLY zero outside the square
    // [wr0, wr0 + wn) x [wc0, wc0 + wn) and the convolution result fits inside the domain, so
    // the step is a plain linear convolution of that window on a torus N >= wn + 2m that may be
    // much smaller than the full one: nothing wraps, nothing folds, no flag can trip.  The
    // result lands in [wr0 - m, wr0 + wn + m) x [wc0 - m, wc0 + wn + m); everything else stays zero.
    int win, wr0, wc0, wn;
};

// Layout of the transposed spectra Yt / Wt / Krt: spectral columns are grouped in
// tiles of PKB_CB; element (column k, row r) lives at ((k / CB) * ld + r) * CB + k % CB.
// The row kernels (which own a row pair and sweep over k) then touch CB * 16 B
// contiguous bytes per row and tile -- whole 128-byte lines at CB = 8, with the two
// rows of a pair adjacent -- instead of one 32-byte sector every ld * 16 B, which
// HBM serves at a fraction of its streaming rate.  The column kernel (which owns a
// column and sweeps over r) walks consecutive lines, one 16-byte element per line;
// neighbouring columns are in flight on neighbouring CTAs, so lines fill in L2.
#ifndef PKB_CB
#define PKB_CB 4
#endif
static_assert(PKB_CB == 2 || PKB_CB == 4 || PKB_CB == 8, "PKB_CB must be 2, 4 or 8");
__host__ __device__ __forceinline__ size_t spec_index(int k, int r, int ld) {
    return ((size_t)(k / PKB_CB) * ld + r) * PKB_CB + (k % PKB_CB);
}
__host__ __device__ __forceinline__ size_t spec_size(int ncols, int ld) {
    return (size_t)((ncols + PKB_CB - 1) / PKB_CB) * ld * PKB_CB;
}

struct ChainCtrl {
    int trunc;   // state is zero outside [0,D)^2: only that block is read
    int flag;    // boundary flag of the last step
    int fused;   // the k_rows_inv that produced this state also transformed its interior row pairs for the next step
    // Spectral-resident steps (see "Spectral-resident state" below).
    int spec;    // Shat holds the spectrum of this state and the next whole-torus step may start from it
    int stored;  // message k_cols -> k_rows_inv of one step: k_cols has written the product spectrum to Shat
    int hint;    // the content of this state outside the domain is negligible (<= PKB_SPEC_EPS): worth storing Shat next step
    double eps_sum;   // bound on the deviation from the exact fold mod P accumulated since the state went spectral
    double eps_max;   // largest outside-domain content seen since then
    int er0, er1;     // rows [er0, er1) hold every cell of this state with |value| >= PKB_SPEC_TAU (er1 <= er0: unknown)
    int ec0, ec1;     // the same for columns (measured by support-window steps only; ec1 <= ec0: unknown)
};

// Spectral-resident state.  The reference keeps its chain state in Fourier space and only inverse-transforms
// a copy per day (CalcSol.py:66,189-201).  The exact-P method of this file re-transforms the state every step
// because the fold mod P is a real-space operation.  When NOTHING of consequence lies outside the domain the
// fold moves nothing of consequence: with E = max |content outside [0,D)^2|, skipping it changes a later day
// by at most 2 E per step (the folded cells are convolved with a kernel of unit mass).  So while E stays
// below PKB_SPEC_EPS and the accumulated bound below PKB_SPEC_BUDGET -- both far under the 1e-10 parity bar
// and the 1e-8 flag threshold -- the product spectrum that k_cols forms anyway is kept (Shat) and the next
// step starts from it: no forward row pass, no forward column transform of the state.  The real state is
// still produced (folded, with flag and sums) every day, so the chain drops back to exact steps the moment
// the criterion fails -- decided on the device in the last CTA of k_rows_inv, no host round trip.
// (What lies outside the domain in the in-domain regime is rounding noise of the transforms, ~2e-14 of the state's
// peak on a 4704^2 torus: 1e-14 right after the release, when the peak is ~0.4, 1e-19 a few weeks later.)
#define PKB_SPEC_EPS 1e-13
#define PKB_SPEC_BUDGET 1e-11
// Row window of a spectral-resident step.  The real state of such a step is only produced for output, and most of
// it is (numerically) empty: if every cell of yesterday's state outside the rows [er0, er1) is below PKB_SPEC_TAU in
// magnitude, every cell of today's outside [er0 - m, er1 + m) is too (a convolution with a non-negative kernel of
// unit mass and radius m is a weighted average).  Those rows are neither inverse-transformed nor read by the
// emission; the thresholded output there is zero either way (TAU << 1e-8), and as long as the window stays inside
// the domain rows the content outside the domain is below max(TAU, what the window rows show in their pad columns),
// which is what enters the criterion above.  er0 / er1 are re-measured on every step from the rows that were computed.
#define PKB_SPEC_TAU 1e-15
__device__ __forceinline__ bool spec_row_window(int allow, int spec, int er0, int er1, double eps_sum, double eps_max, int m, int D, int& w0,
                                                int& w1) {
    if (!allow || !spec || er1 <= er0) return false;
    w0 = er0 - m;
    w1 = er1 + m;
    if (w0 < 0 || w1 > D) return false;
    if (eps_sum + 2.0 * fmax(eps_max, PKB_SPEC_TAU) > PKB_SPEC_BUDGET) return false;
    w0 &= ~1;
    w1 = (w1 + 1) & ~1;
    return true;
}

// Geometry of a step whose SOURCE state is truncated (ChainCtrl::trunc: zero outside [0,D)^2).  Its linear
// convolution with a radius-m kernel spans D + 2m cells only, so such a step can run on a torus
// N_t >= D + 2m < P + 2m.  Whether the source is truncated is only known on the device, so every
// whole-torus launch carries both geometries and picks one at its start (uniformly over the grid).
struct TruncGeom {
    int N, Nc, ldW;          // torus side, spectral columns, Wt leading dimension (0: no smaller torus, use the step's own)
    int cols_kb;             // k_cols geometry of that plan at the launch's CTA size
};

struct StepMeta {   // one per emitted solution
    double padmax, ksum, add, padabs;   // padabs: max |value| outside the domain
    long long kcnt;
    int flag;
    int spec;    // this step started from the stored spectrum (spectral-resident step)
    int wr0, wr1;   // rows [wr0, wr1) of the state were computed, the others are below PKB_SPEC_TAU (wr1 <= wr0: all rows)
    int wc0, wc1;   // ... and columns [wc0, wc1) (support-window steps; wc1 <= wc0: all columns)
    int er0, er1, ec0, ec1;   // measured extent of the cells >= PKB_SPEC_TAU of this state (copy of ChainCtrl's)
};

// ---------------------------------------------------------------------------
// The three FFT kernels are PERSISTENT: the grid is (resident CTAs per SM) x
// (SM count) and every CTA loops over its jobs, so the base twiddles are staged
// in shared memory once per CTA and k_cols' per-CTA scratch stays L2 resident.
//
// Dynamic shared memory of each: fft_smem_bytes(plan) = (N + ntw) complex.

// Hermitian split of the transform of z = a + i b (a, b real rows):
//   A_k = (Z_k + conj(Z_{N-k})) / 2,  B_k = (Z_k - conj(Z_{N-k})) / (2i)
__device__ __forceinline__ void hermitian_split(cplx zk, cplx zn, cplx& A, cplx& B) {
    A = cmake(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
    B = cmake(0.5 * (zk.y + zn.y), 0.5 * (zn.x - zk.x));
}

#ifndef PKB_UNPACK_U
#define PKB_UNPACK_U 2   // column PAIRS in flight per thread in the pack / unpack loops (deeper did not help: not latency bound)
#endif
#ifndef PKB_PREFETCH
#define PKB_PREFETCH 0    // bit 0: rows_fwd, bit 1: rows_inv prefetch their next job into L2 (measured with 3 CTAs per SM: neither pays)
#endif
// column PAIRS in flight per thread in the pack / unpack loops (covers N <= 5120 at 256 threads in one sweep)

// x holds the digit-reversed transform of two packed real rows a (row r) and b (row r + 1);
// write their half spectra A_k, B_k into the tiled transposed array dst.  Each thread
// handles two adjacent columns (k, k + 1) so that every store is 32 bytes.
// The caller has NOT yet synchronised after the last forward stage: the first sweep's
// gather indices are fetched (L2 latency) before the barrier.
__device__ __forceinline__ void unpack_store(const cplx* x, const FftPlan& plan, int Nc, cplx* __restrict__ dst, int ld, int r, bool two,
                                             int tid, int T) {
    const int npair = (Nc + 1) / 2;
    bool synced = false;
    for (int j0 = tid; j0 < npair || !synced; j0 += PKB_UNPACK_U * T) {
        int4 pr[PKB_UNPACK_U];
        int col[PKB_UNPACK_U];       // column pair handled by this slot
#pragma unroll
        for (int u = 0; u < PKB_UNPACK_U; ++u) {
            const int j = j0 + u * T;
            pr[u] = j < npair ? __ldg(plan.spair + j) : make_int4(0, 0, 0, 0);
            col[u] = j < npair ? __ldg(plan.slot + j) : 0;
        }
        if (!synced) {
            __syncthreads();
            synced = true;
        }
#pragma unroll
        for (int u = 0; u < PKB_UNPACK_U; ++u) {
            const int j = j0 + u * T;
            if (j < npair) {
                const cplx zk0 = x[pr[u].x], zn0 = x[pr[u].y], zk1 = x[pr[u].z], zn1 = x[pr[u].w];
                cplx A0, B0, A1, B1;
                hermitian_split(zk0, zn0, A0, B0);
                hermitian_split(zk1, zn1, A1, B1);
                cplx* o = dst + spec_index(2 * col[u], r, ld);
                st_pair(o, A0, A1);
                if (two) st_pair(o + PKB_CB, B0, B1);
            }
        }
    }
}

// grid = persistent, block = T
// pre_m >= 0: the k_rows_inv that produced S (filter radius pre_m) has already transformed the
// complete interior row pairs of the new state into Yt (fused_pair below); unless that state turned
// out flagged -- then its truncated form must be transformed from scratch -- only the border
// pairs are left for this kernel.
__host__ __device__ __forceinline__ bool fused_pair(int r0, int P, int m) {
    return r0 >= m + (m & 1) && r0 + 1 < P - m;       // both rows interior, pair on an even row
}
// the fusion needs room for the reduction scratch in the (zero) tail of the transform buffer
__host__ __device__ __forceinline__ bool rows_fusable(int N, int P) { return N - 48 >= P; }
__global__ void PKB_ROWS_LB k_rows_fwd(const double* __restrict__ S, ChainDims d, const ChainCtrl* ctrl,
                                       cplx* __restrict__ Yt, FftPlan plan, int pre_m, TruncGeom tg, FftPlan plan_t, int spec_try) {
    if (spec_try && ctrl->spec) return;      // spectral-resident step: k_cols starts from Shat, nothing to transform
    if (tg.N && ctrl->trunc) {      // truncated source: the smaller torus (TruncGeom)
        d.N = tg.N; d.Nc = tg.Nc; d.ldW = tg.ldW;
        plan = plan_t;
    }
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    cplx* tws = x + plan.N;
    const int tid = threadIdx.x, T = blockDim.x;
    fft_load_twiddles(tws, plan, tid, T);
    __syncthreads();
    const int lim = d.win ? d.wn : (ctrl->trunc ? d.D : d.P);
    const int njobs = (lim + 1) / 2;
    if (d.win) S += (size_t)d.wr0 * d.ldS + d.wc0;     // rows / columns below are relative to the window
    const bool skip_interior = pre_m >= 0 && !d.win && !ctrl->trunc && ctrl->fused;
    // border pairs only: [0, lo/2) and [hi_job, njobs), the interior pairs in between are done
    const int lo_job = skip_interior ? (pre_m + (pre_m & 1)) / 2 : njobs;
    const int hi_job = skip_interior ? (d.P - pre_m) / 2 : njobs;       // first pair with r0 + 1 >= P - pre_m
    const int nwork = skip_interior ? lo_job + (njobs - hi_job) : njobs;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int job = (skip_interior && w >= lo_job) ? hi_job + (w - lo_job) : w;
        const int r0 = 2 * job;
        const bool two = r0 + 1 < lim;
        const double* s0 = S + (size_t)r0 * d.ldS;
        const double* s1 = s0 + (two ? d.ldS : 0);
        auto ld = [&](int j) -> cplx {
            if (j >= lim) return cmake(0.0, 0.0);
            return cmake(s0[j], two ? s1[j] : 0.0);
        };
        {   // next job's two rows -> L2 while this one is transformed
            const int nr0 = 2 * (job + (int)gridDim.x);
            if ((PKB_PREFETCH & 1) && !skip_interior && nr0 < lim) {
                const char* nxt = reinterpret_cast<const char*>(S + (size_t)nr0 * d.ldS);
                const int nbytes = (nr0 + 1 < lim ? 2 : 1) * d.ldS * (int)sizeof(double);
                for (int o = tid * 128; o < nbytes; o += T * 128) prefetch_l2(nxt + o);
            }
        }
        fft_forward_from(x, tws, plan, tid, T, ld, false);
        unpack_store(x, plan, d.Nc, Yt, d.ldY, r0, two, tid, T);
        __syncthreads();
    }
}

// Row spectra of the wrap-shifted kernel (CalcSol.py:58-64 on the N torus).
// K: dense (Wk x Wk) window centred on the release cell, support radius m.
// Krt column k, row q: q = dy for dy in [0,m], q = dy + 2m+1 for dy in [-m,-1].
// One job = one row pair (q0, q0 + 1), q0 = 2 * job.
__device__ __forceinline__ void kernel_rows_job(cplx* x, const cplx* tws, const double* __restrict__ K, int Wk, int m, const ChainDims& d,
                                                cplx* __restrict__ Krt, const FftPlan& plan, int job, int tid, int T) {
    const int nq = 2 * m + 1;
    const int ck = Wk / 2;
    const int N = d.N;
    const int q0 = 2 * job;
    const bool two = q0 + 1 < nq;
    const int dy0 = q0 <= m ? q0 : q0 - nq;
    const int dy1 = !two ? 0 : (q0 + 1 <= m ? q0 + 1 : q0 + 1 - nq);
    const double* k0 = K + (size_t)(ck + dy0) * Wk + ck;
    const double* k1 = K + (size_t)(ck + dy1) * Wk + ck;
    auto ld = [&](int j) -> cplx {
        int dx;
        if (j <= m) dx = j;
        else if (j >= N - m) dx = j - N;
        else return cmake(0.0, 0.0);
        return cmake(k0[dx], two ? k1[dx] : 0.0);
    };
    fft_forward_from(x, tws, plan, tid, T, ld, false);
    unpack_store(x, plan, d.Nc, Krt, d.ldK, q0, two, tid, T);
    __syncthreads();
}

// grid = persistent over the m+1 row pairs, block = T
__global__ void PKB_ROWS_LB k_kernel_rows(const double* __restrict__ K, int Wk, int m, ChainDims d, cplx* __restrict__ Krt,
                                          FftPlan plan) {
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    cplx* tws = x + plan.N;
    const int tid = threadIdx.x, T = blockDim.x;
    fft_load_twiddles(tws, plan, tid, T);
    __syncthreads();
    for (int job = blockIdx.x; job <= m; job += gridDim.x) kernel_rows_job(x, tws, K, Wk, m, d, Krt, plan, job, tid, T);
}

// Row spectra of the kernels of a whole block of days in ONE launch (the fused
// solve: phase 1 has left every day's kernel on the device, so none of this work
// needs to sit on the chain's critical path).  Day i: window K0 + i * kstride,
// radius b.m[i], spectra to Krt0 + i * krt_stride; jobs [b.job0[i], b.job0[i+1]).
#define PKB_KR_MAXD 64
struct KrBatch {
    int nd;
    int job0[PKB_KR_MAXD + 1];
    int m[PKB_KR_MAXD];
};
__global__ void PKB_ROWS_LB k_kernel_rows_batch(const double* __restrict__ K0, size_t kstride, int Wk, KrBatch b, ChainDims d,
                                                cplx* __restrict__ Krt0, size_t krt_stride, FftPlan plan) {
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    cplx* tws = x + plan.N;
    const int tid = threadIdx.x, T = blockDim.x;
    fft_load_twiddles(tws, plan, tid, T);
    __syncthreads();
    const int total = b.job0[b.nd];
    int day = 0;
    for (int job = blockIdx.x; job < total; job += gridDim.x) {
        while (job >= b.job0[day + 1]) ++day;      // jobs of a CTA increase monotonically
        kernel_rows_job(x, tws, K0 + (size_t)day * kstride, Wk, b.m[day], d, Krt0 + (size_t)day * krt_stride, plan,
                        job - b.job0[day], tid, T);
    }
}

// ---------------------------------------------------------------------------
// Last forward stage of a column fused with the spectral product and the first
// inverse stage.  Thread t owns the final-stage blocks t, t + T, ... (R_last
// contiguous points each).  Phase 0 (filter column): DFT of each block, result
// parked in the CTA's global scratch (L2 resident; each thread re-reads only
// what it wrote, laid out [kb][q][tid] so that both directions coalesce).
// Phase 1 (state column): DFT, product with the parked filter spectrum,
// inverse DFT, back to the same shared-memory slots.
// Spectral-resident steps (ChainCtrl::spec): the product spectrum is also written to the column's slice of Shat
// (phase 3), and a later step reads it from there instead of transforming the state column (phase 2).  Shat uses
// the parking layout [kb][q][tid] -- only this loop ever touches it, with the same plan and CTA size.
template <int RL>
__device__ __forceinline__ void cols_final(cplx* x, cplx* __restrict__ scr, cplx* __restrict__ hcol, int tid, int T, int nbl, int phase) {
    int slot = tid;
#pragma unroll 1
    for (int j = tid; j < nbl; j += T, slot += RL * T) {
        cplx v[RL];
        if (phase == 2) {
            cplx kf[RL];
#pragma unroll
            for (int q = 0; q < RL; ++q) kf[q] = scr[slot + q * T];
#pragma unroll
            for (int q = 0; q < RL; ++q) v[q] = hcol[slot + q * T];
#pragma unroll
            for (int q = 0; q < RL; ++q) v[q] = cmul_f(v[q], kf[q]);
#pragma unroll
            for (int q = 0; q < RL; ++q) hcol[slot + q * T] = v[q];
            idft<RL>(v);
#pragma unroll
            for (int q = 0; q < RL; ++q) x[j * RL + q] = v[q];
        } else if (phase == 0) {
#pragma unroll
            for (int q = 0; q < RL; ++q) v[q] = x[j * RL + q];
            dft<RL>(v);
#pragma unroll
            for (int q = 0; q < RL; ++q) scr[slot + q * T] = v[q];
        } else {
            cplx kf[RL];
#pragma unroll
            for (int q = 0; q < RL; ++q) kf[q] = scr[slot + q * T];
#pragma unroll
            for (int q = 0; q < RL; ++q) v[q] = x[j * RL + q];
            dft<RL>(v);
#pragma unroll
            for (int q = 0; q < RL; ++q) v[q] = cmul_f(v[q], kf[q]);
            if (phase == 3) {
#pragma unroll
                for (int q = 0; q < RL; ++q) hcol[slot + q * T] = v[q];
            }
            idft<RL>(v);
#pragma unroll
            for (int q = 0; q < RL; ++q) x[j * RL + q] = v[q];
        }
    }
}

// grid = persistent, block = plan.cols_threads.
// scr: gridDim.x slices of plan.cols_kb * R_last * blockDim.x complex.
// Per spectral column: forward FFT of the filter column (inputs straight from
// Krt, only 2m+1 of them non-zero), forward FFT of the state column (inputs
// straight from Yt), product, inverse FFT, rows needed by the fold straight to Wt.
// Shat (optional): spectral-resident state, gridDim-independent slices of hstride complex per column.
__global__ void PKB_COLS_LB k_cols(const cplx* __restrict__ Yt, const cplx* __restrict__ Krt, int m, ChainDims d,
                                   ChainCtrl* ctrl, cplx* __restrict__ Wt, cplx* __restrict__ scr, FftPlan plan,
                                   TruncGeom tg, FftPlan plan_t, const cplx* __restrict__ Krt_t, cplx* __restrict__ Shat, size_t hstride,
                                   int rowwin_ok) {
    const int c_trunc = ctrl->trunc, c_spec = ctrl->spec, c_hint = ctrl->hint;
    const int c_er0 = ctrl->er0, c_er1 = ctrl->er1;
    const double c_esum = ctrl->eps_sum, c_emax = ctrl->eps_max;
    const bool tr = tg.N && c_trunc;          // truncated source on its smaller torus (TruncGeom)
    // start from the stored spectrum / keep the product spectrum for the next step (uniform over the grid; `stored`
    // tells this step's k_rows_inv, which decides whether the next step may start from it)
    const bool use_spec = Shat != nullptr && c_spec && !tr && !d.win;
    const bool store = Shat != nullptr && !tr && !d.win && (use_spec || c_hint);
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl->stored = store ? 1 : 0;
    // spectral-resident step: only the rows of the window are needed downstream (spec_row_window)
    int w0 = 0, w1 = 0;
    const bool rowwin = use_spec && spec_row_window(rowwin_ok, c_spec, c_er0, c_er1, c_esum, c_emax, m, d.D, w0, w1);
    if (tr) {
        d.N = tg.N; d.Nc = tg.Nc; d.ldW = tg.ldW;
        plan = plan_t;
        plan.cols_kb = tg.cols_kb;
        Krt = Krt_t;
    }
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    cplx* tws = x + plan.N;
    const int tid = threadIdx.x, T = blockDim.x;
    fft_load_twiddles(tws, plan, tid, T);
    __syncthreads();
    const int lim = d.win ? d.wn : (c_trunc ? d.D : d.P);
    const int N = d.N, nq = 2 * m + 1;
    const int L = plan.nstage, R0 = plan_radix(plan, 0), RL = plan_radix(plan, L - 1), nbl = N / RL;
    const int hi = (d.win ? d.wn : (tr ? d.D : d.P)) + m;   // rows [0, extent + m) and [N-m, N) are needed downstream
    cplx* myscr = scr + (size_t)blockIdx.x * ((size_t)plan.cols_kb * RL * T);
    const int off_last = plan.ntw - 1;   // table offset of the last stage (one entry)
    for (int c = blockIdx.x; c < d.Nc; c += gridDim.x) {
        // row i of this column lives at col[i * PKB_CB]
        const cplx* ycol = Yt + spec_index(c, 0, d.ldY);
        const cplx* kcol = Krt + spec_index(c, 0, d.ldK);
        cplx* wcol = Wt + spec_index(c, 0, d.ldW);
        cplx* hcol = store ? Shat + (size_t)c * hstride : nullptr;
        auto ld_filter = [&](int i) -> cplx {
            if (i <= m) return kcol[(size_t)i * PKB_CB];
            if (i >= N - m) return kcol[(size_t)(i - (N - nq)) * PKB_CB];
            return cmake(0.0, 0.0);
        };
        auto ld_state = [&](int i) -> cplx { return i < lim ? ycol[(size_t)i * PKB_CB] : cmake(0.0, 0.0); };
        auto st_out = [&](int i, cplx v) {
            if (rowwin ? (i >= w0 && i < w1) : (i < hi || i >= N - m)) wcol[(size_t)i * PKB_CB] = v;
        };
        for (int phase = 0; phase < 2; ++phase) {
            auto ld = [&](int i) -> cplx { return phase ? ld_state(i) : ld_filter(i); };
            if (phase && use_spec) {
                // the state column's spectrum is in Shat: nothing to transform
            } else if (L == 1) {
                for (int i = tid; i < N; i += T) x[i] = ld(i);
                __syncthreads();
            } else {
                fft_stage_dispatch<false, true>(R0, N, N, plan.tw0, tid, T, ld, SmemStore{x});
                __syncthreads();
                fft_fwd_stages(x, tws, plan, 1, L - 1, N / R0, 0, tid, T);
            }
            const int fmode = !phase ? 0 : (use_spec ? 2 : (store ? 3 : 1));
#define PKB_CALL_(RR) cols_final<RR>(x, myscr, hcol, tid, T, nbl, fmode)
            PKB_RADIX_SWITCH(RL, PKB_CALL_)
#undef PKB_CALL_
            __syncthreads();
        }
        if (L == 1) {
            for (int i = tid; i < N; i += T) st_out(i, x[i]);
        } else {
            fft_inv_stages(x, tws, plan, L - 1, 1, RL, off_last, tid, T);
            fft_stage_dispatch<true, true>(R0, N, N, plan.tw0, tid, T, SmemLoad{x}, st_out);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Per-row statistics written by k_rows_inv, reduced by k_step_finalize.
struct RowStats {
    double padmax, ksum, padabs;
    int kcnt;
    int has_e;      // some cell of the row (any column of the torus) is >= PKB_SPEC_TAU in magnitude
};
// has_e travels through the block reduction inside the kept-count sum: every thread that saw such a cell adds 2^20
// (a row has far fewer than 2^20 cells, and the sums are exact in fp64)
#define PKB_HAS_E_UNIT 1048576.0

// Block reduction of 8 per-thread values: entries with (i & 3) == 0 or 3 by max, the
// others by sum.  Result valid in thread i (i < 8) of warp 0.  red: PKB_RED_DOUBLES doubles
// (8 stats x up to 8 warps, then 8 results); the FFT kernels alias it onto the start of their
// (then idle) transform buffer so that they carry no static shared memory.
#define PKB_RED_DOUBLES 72
__device__ __forceinline__ double block_reduce8(double (&v)[8], double* red, int tid, int T) {
    const int lane = tid & 31, wid = tid >> 5, nw = (T + 31) >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool ismax = (i & 3) == 0 || (i & 3) == 3;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double t = __shfl_down_sync(0xffffffffu, v[i], o);
            v[i] = ismax ? fmax(v[i], t) : v[i] + t;
        }
        if (lane == 0) red[i * 8 + wid] = v[i];
    }
    __syncthreads();
    double r = 0.0;
    if (tid < 8) {
        const bool ismax = (tid & 3) == 0 || (tid & 3) == 3;
        r = red[tid * 8];
        for (int w = 1; w < nw; ++w) r = ismax ? fmax(r, red[tid * 8 + w]) : r + red[tid * 8 + w];
    }
    return r;
}

// What the finalising CTA needs to know about the step's SOURCE state and this step's k_cols to decide
// whether the NEXT step may start from Shat (read at kernel start, before anything is overwritten).
struct SpecIn {
    int stored;        // k_cols of this step wrote the product spectrum
    int was_spec;      // this step itself started from Shat
    double eps_sum, eps_max;
    int rowwin, w0, w1;   // only the rows [w0, w1) were computed (spec_row_window, or a support-window step)
    int c0, c1;           // ... and only the columns [c0, c1) (support-window step; c1 <= c0: all)
    int* colflag;         // optional [P]: columns in which this step saw a cell >= PKB_SPEC_TAU (cleared again here)
    int* host_box;        // optional, MAPPED PINNED HOST memory [6]: the measured extent (er0, er1, ec0, ec1), then a ticket the
    int ticket;           //   host spins on -- written straight from the kernel so that the host learns it microseconds later
};

// Flag / kept sum / kept count / largest outside-domain magnitude of a state from its per-row statistics
// -> control block and step meta (CalcSol.py:36-37, 134-135).  Whole CTA; fixed
// summation order for a given block size.  red: 8 * 32 doubles.  flag_thresh: 1e-8 (CalcSol.py:37), or the
// caller's negval on the cuda_lib.get_cursol path (cuda_lib.py:117-130).
__device__ __forceinline__ void step_finalize_block(const RowStats* rstat, const ChainDims& d, ChainCtrl* ctrl,
                                                    StepMeta* __restrict__ meta, int apply_trunc, double* red, int tid, int T,
                                                    SpecIn si, double flag_thresh) {
    // st[0] pad max, st[1] kept sum, st[2] kept count, st[3] max |v| outside the domain,
    // st[4] -(first row with a cell >= PKB_SPEC_TAU), st[7] last such row
    double st[8] = {-INFINITY, 0.0, 0.0, 0.0, -INFINITY, 0.0, 0.0, -INFINITY};
    // contiguous chunks per thread so the summation order over rows is fixed; a row-windowed step only has
    // statistics for its window rows (the others hold nothing above PKB_SPEC_TAU)
    const int lo = si.rowwin ? si.w0 : 0, hi = si.rowwin ? si.w1 : d.P;
    const int chunk = (hi - lo + T - 1) / T;
    const volatile RowStats* rv = rstat;
    for (int r = lo + tid * chunk; r < hi && r < lo + (tid + 1) * chunk; ++r) {
        const double pm = rv[r].padmax, ks = rv[r].ksum, pa = rv[r].padabs;
        const int kc = rv[r].kcnt;
        st[0] = fmax(st[0], pm);
        st[3] = fmax(st[3], pa);
        if (r < d.D) { st[1] += ks; st[2] += (double)kc; }
        if (rv[r].has_e) { st[4] = fmax(st[4], -(double)r); st[7] = fmax(st[7], (double)r); }
    }
    __syncthreads();
    const double rr = block_reduce8(st, red, tid, T);
    if (tid < 8) red[64 + tid] = rr;
    __syncthreads();
    if (tid == 0) {
        const double bp = si.rowwin ? fmax(red[64], -PKB_SPEC_TAU) : red[64], bs = red[65], bc = red[66];
        const double ba = si.rowwin ? fmax(red[67], PKB_SPEC_TAU) : red[67];
        const int flag = bp > flag_thresh ? 1 : 0;   // CalcSol.py:36-37
        meta->padmax = bp; meta->ksum = bs; meta->kcnt = (long long)bc; meta->padabs = ba;
        meta->add = (1.0 - bs) / bc;                 // CalcSol.py:135
        meta->flag = flag;
        meta->spec = si.was_spec;
        meta->wr0 = si.rowwin ? si.w0 : 0;
        meta->wr1 = si.rowwin ? si.w1 : 0;
        meta->wc0 = si.c1 > si.c0 ? si.c0 : 0;
        meta->wc1 = si.c1 > si.c0 ? si.c1 : 0;
        ctrl->flag = flag;
        meta->er0 = ctrl->er0 = red[71] >= 0.0 ? (int)(-red[68]) : 0;
        meta->er1 = ctrl->er1 = red[71] >= 0.0 ? (int)red[71] + 1 : 0;
        // a fresh convolution result is a full P x P state; it becomes a
        // truncated one only where the caller applies CalcSol.py:200-201
        ctrl->trunc = apply_trunc ? flag : 0;
        // spectral-resident steps: see PKB_SPEC_EPS
        const bool small = ba <= PKB_SPEC_EPS;
        const double emax = fmax(si.was_spec ? si.eps_max : 0.0, ba);
        const double esum = (si.was_spec ? si.eps_sum : 0.0) + 2.0 * emax;
        const bool spec = si.stored && small && !flag && esum <= PKB_SPEC_BUDGET;
        ctrl->hint = small ? 1 : 0;
        ctrl->spec = spec ? 1 : 0;
        ctrl->stored = 0;
        ctrl->eps_sum = spec ? esum : 0.0;
        ctrl->eps_max = spec ? emax : 0.0;
    }
    // column extent of the cells >= PKB_SPEC_TAU (support-window steps mark the columns in colflag)
    if (si.colflag) {
        __syncthreads();
        double cs[8] = {-INFINITY, 0.0, 0.0, -INFINITY, -INFINITY, 0.0, 0.0, -INFINITY};
        for (int c = tid; c < d.P; c += T)
            if (si.colflag[c]) {
                cs[0] = fmax(cs[0], -(double)c);
                cs[3] = fmax(cs[3], (double)c);
                si.colflag[c] = 0;
            }
        const double cr = block_reduce8(cs, red, tid, T);
        if (tid < 4) red[64 + tid] = cr;
        __syncthreads();
        if (tid == 0) {
            meta->ec0 = ctrl->ec0 = red[67] >= 0.0 ? (int)(-red[64]) : 0;
            meta->ec1 = ctrl->ec1 = red[67] >= 0.0 ? (int)red[67] + 1 : 0;
            if (si.host_box) {
                volatile int* hb = si.host_box;
                hb[0] = ctrl->er0; hb[1] = ctrl->er1; hb[2] = ctrl->ec0; hb[3] = ctrl->ec1;
                __threadfence_system();
                hb[4] = si.ticket;
            }
        }
    } else if (tid == 0) {
        meta->ec0 = ctrl->ec0 = 0;
        meta->ec1 = ctrl->ec1 = 0;
    }
}

// grid = 1, block = 256 (stencil path and re-thresholding; the FFT path finalises
// in the last CTA of k_rows_inv)
__global__ void k_step_finalize(const RowStats* __restrict__ rstat, ChainDims d, ChainCtrl* ctrl, StepMeta* __restrict__ meta,
                                int apply_trunc, double flag_thresh) {
    PKB_SHARED(double, red, PKB_RED_DOUBLES);
    SpecIn si = {0, 0, 0.0, 0.0, 0, 0, 0, 0, 0, nullptr, nullptr, 0};
    step_finalize_block(rstat, d, ctrl, meta, apply_trunc, red, threadIdx.x, blockDim.x, si, flag_thresh);
}

// number of inverse-row jobs: 2m fold jobs, an unpaired row m if m is odd, then row pairs
__host__ __device__ __forceinline__ int rows_inv_jobs(int P, int m) {
    if (P - 2 * m <= 0) return 2 * m;
    const int lo = m + (m & 1);
    const int rest = P - m - lo;
    return 2 * m + (m & 1) + (rest > 0 ? (rest + 1) / 2 : 0);
}

// rows of job `job`: inputs (ra, rb) of Wt, outputs (out_a, out_b) of the state, fold flag
__device__ __forceinline__ void rows_inv_decode(int job, int m, int P, int N, int& ra, int& rb, int& out_a, int& out_b, bool& fold) {
    if (job < m) { fold = true; out_a = job; out_b = -1; ra = job; rb = job + P; }
    else if (job < 2 * m) { const int t = job - m; fold = true; out_a = P - m + t; out_b = -1; ra = out_a; rb = N - m + t; }
    else {
        // interior rows [m, P-m): pairs start on even rows
        int j = job - 2 * m, r;
        if (m & 1) { r = j == 0 ? m : m + 1 + 2 * (j - 1); out_b = (j > 0 && r + 1 < P - m) ? r + 1 : -1; }
        else { r = m + 2 * j; out_b = (r + 1 < P - m) ? r + 1 : -1; }
        fold = false; out_a = r; ra = r; rb = out_b;
    }
}

// Truncated source on its own torus N_t (TruncGeom): linear-convolution row j lives in Wt row j for
// j in [0, E), E = D + m, and in row N_t + j for j in [-m, 0); nothing else is non-zero.  Output row r of
// the P torus = lin[r] (r < E) + lin[r - P] (r >= Lo, Lo = P - m).  Rows [0, min(E, Lo)) and [max(E, Lo), P)
// have one source each and are paired two per transform; in between the rows either have both sources
// (E > Lo: "fold" jobs, one row per transform) or none (E <= Lo: "zero" jobs, two rows each, no transform).
__host__ __device__ __forceinline__ int rows_inv_jobs_trunc(int P, int D, int m) {
    const int E = D + m, Lo = P - m;
    const int a1 = E < Lo ? E : Lo, b1 = E < Lo ? Lo : E;
    return (a1 + 1) / 2 + (E > Lo ? b1 - a1 : (b1 - a1 + 1) / 2) + (P - b1 + 1) / 2;
}
// kind: 0 pair, 1 fold, 2 zero
__device__ __forceinline__ void rows_inv_decode_trunc(int job, int m, int P, int D, int Nt, int& ra, int& rb, int& out_a, int& out_b, int& kind) {
    const int E = D + m, Lo = P - m;
    const int a1 = E < Lo ? E : Lo, b1 = E < Lo ? Lo : E;
    const int nA = (a1 + 1) / 2, nM = E > Lo ? b1 - a1 : (b1 - a1 + 1) / 2;
    if (job < nA) {
        kind = 0; out_a = 2 * job; out_b = out_a + 1 < a1 ? out_a + 1 : -1; ra = out_a; rb = out_b;
    } else if (job < nA + nM) {
        const int j = job - nA;
        if (E > Lo) { kind = 1; out_a = a1 + j; out_b = -1; ra = out_a; rb = Nt - (P - out_a); }
        else { kind = 2; out_a = a1 + 2 * j; out_b = out_a + 1 < b1 ? out_a + 1 : -1; ra = rb = -1; }
    } else {
        const int j = job - nA - nM;
        kind = 0; out_a = b1 + 2 * j; out_b = out_a + 1 < P ? out_a + 1 : -1;
        ra = Nt - (P - out_a); rb = out_b >= 0 ? Nt - (P - out_b) : -1;
    }
}

// Output loop of an inverse-row job: the new state row(s) from the transform buffer x, with the per-row statistics
// (st[0..3]: row a -- pad max, kept sum, kept count, max |pad|; st[4..7]: row b) and the "holds a cell >= PKB_SPEC_TAU" flags.
// One instantiation per geometry so that the loop carries no mode tests: MODE 0 whole torus (columns already folded in
// place), 1 support window (the result lies inside the domain: no pad cells, no fold; the columns holding a cell >= TAU are
// collected in a per-thread bit mask -- bit k for column tid + k T -- instead of one global store per cell), 2 truncated
// source (columns fold like the rows, straight from the transform).  Same arithmetic and order as one generic loop.
template <int MODE>
__device__ __forceinline__ void rows_inv_emit(const cplx* x, double* __restrict__ dst_a, double* __restrict__ dst_b, bool has_b, bool fold, bool zjob,
                                              int ncols, int m, int N, int P, int E, int Lo, bool pad_a, bool pad_b, int Dc, double scale,
                                              double negval, double (&st)[8], bool& e_a, bool& e_b, unsigned& cmask, int* colflag_far, int tid, int T) {
    const cplx zero = cmake(0.0, 0.0);
    int kk = 0;
    for (int c = tid; c < ncols; c += T, ++kk) {
        cplx z;
        if (MODE == 2) {
            z = (c < E && !zjob) ? x[c] : zero;
            if (c >= Lo && !zjob) z = cadd(z, x[N - (P - c)]);
        } else if (MODE == 1) {
            z = x[c < m ? c - m + N : c - m];
        } else {
            z = x[c];
        }
        const double va = ((MODE != 1 && fold) ? z.x + z.y : z.x) * scale;
        dst_a[c] = va;
        const bool ea1 = fabs(va) >= PKB_SPEC_TAU;
        e_a |= ea1;
        if (MODE != 1 && (pad_a || c >= Dc)) { st[0] = fmax(st[0], va); st[3] = fmax(st[3], fabs(va)); }
        else if (!(va < negval)) { st[1] += va; st[2] += 1.0; }
        bool eb1 = false;
        if (has_b) {
            const double vb = z.y * scale;
            dst_b[c] = vb;
            eb1 = fabs(vb) >= PKB_SPEC_TAU;
            e_b |= eb1;
            if (MODE != 1 && (pad_b || c >= Dc)) { st[4] = fmax(st[4], vb); st[7] = fmax(st[7], fabs(vb)); }
            else if (!(vb < negval)) { st[5] += vb; st[6] += 1.0; }
        }
        if (MODE == 1 && (ea1 || eb1)) {
            if (kk < 32) cmask |= 1u << kk;
            else if (colflag_far) colflag_far[c] = 1;      // (more than 32 columns per thread: straight to the flags)
        }
    }
}

// grid = persistent over rows_inv_jobs(P, m) jobs, block = T
// A job is one inverse transform.  "Pair" jobs carry two interior output rows as
// real and imaginary part; "fold" jobs carry the two linear-convolution rows that
// fold onto the same output row mod P (their sum is re + im).
// ctrl (written by the CTA that finishes last) and src_ctrl (read by every CTA at its start) are the SAME object on
// main-chain steps: neither is __restrict__, and the reads below come before this CTA's contribution to `done`, so
// they are ordered before the finalising CTA's writes.
__global__ void PKB_ROWS_LB k_rows_inv(const cplx* __restrict__ Wt, int m, ChainDims d, double* __restrict__ Sout,
                                       RowStats* __restrict__ rstat, double negval, FftPlan plan, int* __restrict__ done,
                                       ChainCtrl* ctrl, StepMeta* __restrict__ meta, int apply_trunc,
                                       cplx* __restrict__ Yt_next, const ChainCtrl* src_ctrl, TruncGeom tg, FftPlan plan_t,
                                       int desc_order, int rowwin_ok, int* __restrict__ colflag, int* host_box, int ticket) {
    const volatile ChainCtrl* sc = src_ctrl;
    SpecIn si = {sc->stored, sc->spec, sc->eps_sum, sc->eps_max, 0, 0, 0, 0, 0, nullptr, nullptr, 0};
    const bool tr = tg.N && !d.win && sc->trunc;      // truncated source on its smaller torus (TruncGeom)
    // row window of a spectral-resident step (the same decision k_cols took from the same control block)
    si.rowwin = (!tr && !d.win && spec_row_window(rowwin_ok, si.was_spec, sc->er0, sc->er1, si.eps_sum, si.eps_max, m, d.D, si.w0, si.w1)) ? 1 : 0;
    if (tr) {
        d.N = tg.N; d.Nc = tg.Nc; d.ldW = tg.ldW;
        plan = plan_t;
    }
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    cplx* tws = x + plan.N;
    // reduction scratch and the "last CTA" flag live at the start of the transform buffer, which is
    // idle whenever they are used (after a job's output loop; fft_smem_bytes() >= 1 KB)
    // (at the END of the buffer when the next step's forward row transform is fused in: the packed
    //  row pair at the start of the buffer is still needed then, and the tail beyond column P is read as zero)
    // (not when this step's k_cols kept the product spectrum: the next step will most likely start from it)
    const bool fuse = Yt_next != nullptr && !d.win && !tr && !si.stored && rows_fusable(plan.N, d.P);
    double* red = reinterpret_cast<double*>(raw) + (fuse ? 2 * (plan.N - 48) : 0);
    int* last = reinterpret_cast<int*>(red + PKB_RED_DOUBLES);
    const int tid = threadIdx.x, T = blockDim.x;
    fft_load_twiddles(tws, plan, tid, T);
    __syncthreads();
    const int P = d.P, N = d.N, D = d.D, Nc = d.Nc;
    const int wout = d.wn + 2 * m;                       // window mode: side of the result
    if (d.win) {
        // only the rows / columns of the result window carry statistics of this step (windows may move and shrink
        // from day to day when they follow the numerical support, pkb200.cu: tau windows)
        si.rowwin = 1; si.w0 = d.wr0 - m; si.w1 = d.wr0 - m + wout;
        si.c0 = d.wc0 - m; si.c1 = d.wc0 - m + wout;
        si.colflag = colflag;
        si.host_box = colflag ? host_box : nullptr;
        si.ticket = ticket;
    }
    const int njobs = d.win ? (wout + 1) / 2 : (tr ? rows_inv_jobs_trunc(P, D, m) : (si.rowwin ? (si.w1 - si.w0) / 2 : rows_inv_jobs(P, m)));
    const double scale = 1.0 / ((double)N * (double)N);
    const cplx zero = cmake(0.0, 0.0);
    const int E = D + m, Lo = P - m;                     // truncated mode: extent of the positive rows / columns, first folded one
    // Whole-torus mode: jobs are taken in DESCENDING order.  The fold jobs (job < 2m) are the short ones --
    // one output row and never a fused forward transform -- so they go last and the partial final round of
    // the persistent grid is made of short jobs instead of the longest ones.
    const bool descending = desc_order && !d.win && !tr && !si.rowwin;
    unsigned cmask = 0;      // support-window steps: columns (tid + k T of the result window) in which this thread saw a cell >= PKB_SPEC_TAU
    for (int it = blockIdx.x; it < njobs; it += gridDim.x) {
        const int job = descending ? njobs - 1 - it : it;
        int ra, rb, out_a, out_b;
        bool fold;
        bool zjob = false;                               // truncated mode: rows without any source (no transform)
        if (tr) {
            int kind;
            rows_inv_decode_trunc(job, m, P, D, N, ra, rb, out_a, out_b, kind);
            fold = kind == 1;
            zjob = kind == 2;
        } else if (d.win) {
            // linear rows j = -m + 2 job and j + 1 (mod N in Wt) -> state rows wr0 + j
            const int ja = 2 * job - m, jb = ja + 1;
            fold = false;
            ra = ja < 0 ? ja + N : ja;
            out_a = d.wr0 + ja;
            if (jb < d.wn + m) { rb = jb < 0 ? jb + N : jb; out_b = d.wr0 + jb; }
            else { rb = -1; out_b = -1; }
        } else if (si.rowwin) {
            // window rows in pairs; what would fold onto them from beyond the torus edge is below PKB_SPEC_TAU
            fold = false;
            out_a = ra = si.w0 + 2 * job;
            out_b = rb = out_a + 1;
        } else {
            rows_inv_decode(job, m, P, N, ra, rb, out_a, out_b, fold);
        }
        {   // next job's Wt rows -> L2 while this one is transformed
            const int nj = job - (int)gridDim.x;        // (descending order)
            if ((PKB_PREFETCH & 2) && !d.win && !tr && !si.rowwin && nj >= 0) {
                int na, nb, oa, ob;
                bool nf;
                rows_inv_decode(nj, m, P, N, na, nb, oa, ob, nf);
                const int ntile = (Nc + PKB_CB - 1) / PKB_CB;
                for (int t = tid; t < ntile; t += T) {
                    prefetch_l2(Wt + spec_index(t * PKB_CB, na, d.ldW));
                    if (nb >= 0) prefetch_l2(Wt + spec_index(t * PKB_CB, nb, d.ldW));
                }
            }
        }
        // scatter the Hermitian pair Z = A + iB into digit-reversed order; each thread
        // handles two adjacent columns (32-byte loads per row)
        const int npair = zjob ? 0 : (Nc + 1) / 2;
        for (int j0 = tid; j0 < npair; j0 += PKB_UNPACK_U * T) {
            int4 pr[PKB_UNPACK_U];
            int col[PKB_UNPACK_U];       // column pair handled by this slot (FftPlan::slot)
            cplx a[2 * PKB_UNPACK_U], b[2 * PKB_UNPACK_U];
#pragma unroll
            for (int u = 0; u < PKB_UNPACK_U; ++u) {
                const int j = j0 + u * T;
                col[u] = j < npair ? __ldg(plan.slot + j) : 0;
            }
#pragma unroll
            for (int u = 0; u < PKB_UNPACK_U; ++u) {
                const int j = j0 + u * T;
                if (j < npair) {
                    pr[u] = __ldg(plan.spair + j);
                    ld_pair(Wt + spec_index(2 * col[u], ra, d.ldW), a[2 * u], a[2 * u + 1]);
                    if (rb >= 0) ld_pair(Wt + spec_index(2 * col[u], rb, d.ldW), b[2 * u], b[2 * u + 1]);
                    else { b[2 * u] = zero; b[2 * u + 1] = zero; }
                } else {
                    pr[u] = make_int4(0, 0, 0, 0);
                    a[2 * u] = a[2 * u + 1] = b[2 * u] = b[2 * u + 1] = zero;
                }
            }
#pragma unroll
            for (int u = 0; u < PKB_UNPACK_U; ++u) {
                const int j = j0 + u * T;
                if (j < npair) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int k = 2 * col[u] + h;
                        if (k >= Nc) break;
                        const cplx av = a[2 * u + h], bv = b[2 * u + h];
                        const int pk = h ? pr[u].z : pr[u].x, pn = h ? pr[u].w : pr[u].y;
                        if (k == 0 || N - k == k) {
                            x[pk] = cmake(av.x, bv.x);        // self-conjugate bins are real
                        } else {
                            x[pk] = cmake(av.x - bv.y, av.y + bv.x);    // A + iB
                            x[pn] = cmake(av.x + bv.y, bv.x - av.y);    // conj(A) + i conj(B)
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (!zjob) fft_inverse_to(x, tws, plan, tid, T, SmemStore{x});
        __syncthreads();
        if (!d.win && !tr) {
            // fold the columns mod P in place (2m <= P, so the two ranges are disjoint)
            for (int c = tid; c < m; c += T) {
                x[c] = cadd(x[c], x[c + P]);
                x[P - m + c] = cadd(x[P - m + c], x[N - m + c]);
            }
            __syncthreads();
        }
        // full mode: output column c <- x[c], c in [0, P); window mode: output column
        // wc0 - m + i <- linear column i - m (mod N), i in [0, wn + 2m)
        const int ncols = d.win ? wout : P;
        const int col0 = d.win ? d.wc0 - m : 0;
        double* dst_a = Sout + (size_t)out_a * d.ldS + col0;
        double* dst_b = Sout + (size_t)(out_b >= 0 ? out_b : out_a) * d.ldS + col0;
        // st[0..3]: row a (pad max, kept sum, kept count, -min); st[4..7]: row b
        double st[8] = {-INFINITY, 0.0, 0.0, 0.0, -INFINITY, 0.0, 0.0, 0.0};
        if (d.win) { st[0] = st[4] = 0.0; }                          // the untouched rest of the row is zero
        const bool pad_a = out_a >= D, pad_b = out_b >= D;
        const int Dc = D - col0;                                     // first pad column, relative to col0
        bool e_a = false, e_b = false;                               // a cell >= PKB_SPEC_TAU in this row (RowStats::has_e)
        int* cfar = si.colflag ? si.colflag + col0 : (int*)nullptr;
        if (tr) rows_inv_emit<2>(x, dst_a, dst_b, out_b >= 0, fold, zjob, ncols, m, N, P, E, Lo, pad_a, pad_b, Dc, scale, negval, st, e_a, e_b, cmask, cfar, tid, T);
        else if (d.win) rows_inv_emit<1>(x, dst_a, dst_b, out_b >= 0, fold, zjob, ncols, m, N, P, E, Lo, pad_a, pad_b, Dc, scale, negval, st, e_a, e_b, cmask, cfar, tid, T);
        else rows_inv_emit<0>(x, dst_a, dst_b, out_b >= 0, fold, zjob, ncols, m, N, P, E, Lo, pad_a, pad_b, Dc, scale, negval, st, e_a, e_b, cmask, cfar, tid, T);
        if (e_a) st[2] += PKB_HAS_E_UNIT;
        if (e_b) st[6] += PKB_HAS_E_UNIT;
        __syncthreads();                               // every thread is done reading x: reuse it as scratch
        const double r = block_reduce8(st, red, tid, T);
        if (tid < 8) red[64 + tid] = r;
        __syncthreads();
        if (tid == 0 || (tid == 1 && out_b >= 0)) {
            const double* q = red + 64 + 4 * tid;
            RowStats rs;
            rs.padmax = q[0]; rs.ksum = q[1]; rs.padabs = q[3];
            rs.has_e = q[2] >= PKB_HAS_E_UNIT ? 1 : 0;
            rs.kcnt = (int)(q[2] - floor(q[2] / PKB_HAS_E_UNIT) * PKB_HAS_E_UNIT);
            rstat[tid ? out_b : out_a] = rs;
        }
        __syncthreads();
        if (fuse && !fold && out_b >= 0 && fused_pair(out_a, P, m)) {
            // The buffer still holds the new row pair packed as re + i im (columns folded, unscaled):
            // exactly the input of the NEXT step's forward row transform.  Scale on load (tail read as zero)
            // and transform here -- the pair never has to be read back from HBM.  (Speculative: if this
            // state ends up flagged, k_rows_fwd redoes every row from the truncated state.)
            auto ld = [&](int c) -> cplx {
                if (c >= P) return zero;
                const cplx z = x[c];
                return cmake(z.x * scale, z.y * scale);
            };
            fft_forward_from(x, tws, plan, tid, T, ld, false);
            unpack_store(x, plan, Nc, Yt_next, d.ldY, out_a, true, tid, T);
            __syncthreads();
        }
    }
    if (si.colflag && cmask) {
        int* cf = si.colflag + (d.wc0 - m);
        for (int kk = 0; kk < 32; ++kk)
            if ((cmask >> kk) & 1u) cf[tid + kk * T] = 1;
    }
    __syncthreads();
    // the CTA that finishes last reduces the row statistics of the whole state
    if (tid == 0) {
        __threadfence();
        last[0] = atomicAdd(done, 1) == (int)gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (last[0]) {
        __threadfence();
        step_finalize_block(rstat, d, ctrl, meta, apply_trunc, red, tid, T, si, 1e-8);
        if (tid == 0) {
            ctrl->fused = fuse ? 1 : 0;
            *done = 0;
        }
    }
}

// S[r][c] = 0 outside the rows [r0, r1) x columns [c0, c1): leaving support-window mode, whose windows may have moved
// and left older days' cells elsewhere in the buffer.  grid = P, block = 256
__global__ void k_zero_outside(double* __restrict__ S, ChainDims d, int r0, int r1, int c0, int c1) {
    const int r = blockIdx.x;
    double* row = S + (size_t)r * d.ldS;
    const bool inside = r >= r0 && r < r1;
    for (int c = threadIdx.x; c < d.P; c += blockDim.x)
        if (!inside || c < c0 || c >= c1) row[c] = 0.0;
}

// get_cursol's "Re-fft" decision taken after the fact (cuda_lib.py:130-136)
__global__ void k_apply_trunc(ChainCtrl* __restrict__ ctrl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) ctrl->trunc = ctrl->flag;
}
// control block of a freshly loaded state: nothing outside the domain (hint), no spectrum kept
__global__ void k_set_ctrl(ChainCtrl* __restrict__ ctrl, int trunc, int flag) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ctrl->trunc = trunc; ctrl->flag = flag; ctrl->fused = 0;
        ctrl->spec = 0; ctrl->stored = 0; ctrl->hint = 1; ctrl->eps_sum = 0.0; ctrl->eps_max = 0.0;
    }
}

// out[r][c] = r_small_vals(S[:D,:D], prob_model) densely (CalcSol.py:112-136)
// grid = D, block = 256
// strict != 0: keep v > negval (cuda_lib.py:117-119) instead of !(v < negval)
// sparse_only != 0: `out` is only an intermediate of the COO / CSR compaction (k_coo_write with the same meta), which reads
// nothing outside the rows and columns this step computed -- the zeros there are not written (a support-window day of the
// 4097^2 solve touches a few hundred columns of its 134 MB grid)
__global__ void k_emit_dense(const double* __restrict__ S, ChainDims d, const StepMeta* __restrict__ meta, double negval, int prob_model,
                             int strict, double* __restrict__ out, int* __restrict__ rownnz, int sparse_only) {
    PKB_SHARED(int, cnt, 1);
    const double add = prob_model ? meta->add : 0.0;
    const int wr0 = meta->wr0, wr1 = meta->wr1;          // row-windowed step: the other rows were not computed (all below PKB_SPEC_TAU)
    const int wc0 = meta->wc1 > meta->wc0 ? meta->wc0 : 0, wc1 = meta->wc1 > meta->wc0 ? meta->wc1 : d.D;      // ... and columns
    // rows blockIdx.x, blockIdx.x + gridDim.x, ...: one CTA per row (grid = D), or a few small persistent
    // CTAs that trickle through the day next to the FFT kernels of the following step (fused solve)
    for (int r = blockIdx.x; r < d.D; r += gridDim.x) {
        const double* src = S + (size_t)r * d.ldS;
        double* dst = out + (size_t)r * d.D;
        if (wr1 > wr0 && (r < wr0 || r >= wr1)) {
            if (!sparse_only)
                for (int c = threadIdx.x; c < d.D; c += blockDim.x) dst[c] = 0.0;
            if (rownnz && threadIdx.x == 0) rownnz[r] = 0;
            continue;
        }
        if (rownnz) {
            __syncthreads();                 // (the previous row's count has been read)
            if (threadIdx.x == 0) cnt[0] = 0;
            __syncthreads();
        }
        int n = 0;
        const int T = blockDim.x;
        const int cbeg = sparse_only ? wc0 : 0, cend = sparse_only ? wc1 : d.D;
        int c = cbeg + threadIdx.x;
        for (; c + 3 * T < cend; c += 4 * T) {          // four independent loads in flight per thread
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (c + u * T >= wc0 && c + u * T < wc1) ? src[c + u * T] : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool keep = strict ? (v[u] > negval) : (v[u] != 0.0 && !(v[u] < negval));
                const double o = keep ? v[u] + add : 0.0;
                dst[c + u * T] = o;
                n += o != 0.0 ? 1 : 0;
            }
        }
        for (; c < cend; c += T) {
            const double v = (c >= wc0 && c < wc1) ? src[c] : 0.0;
            const bool keep = strict ? (v > negval) : (v != 0.0 && !(v < negval));
            const double o = keep ? v + add : 0.0;
            dst[c] = o;
            n += o != 0.0 ? 1 : 0;
        }
        if (rownnz) {       // per-row non-zero count for the COO compaction (saves a pass over the output)
            if (n) atomicAdd(&cnt[0], n);
            __syncthreads();
            if (threadIdx.x == 0) rownnz[r] = cnt[0];
        }
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

// plain copy of the domain block (un-thresholded), grid = D.  meta (optional): rows a row-windowed step did not
// compute (all below PKB_SPEC_TAU) are written as zeros
__global__ void k_copy_domain(const double* __restrict__ S, ChainDims d, double* __restrict__ out, const StepMeta* __restrict__ meta) {
    const int r = blockIdx.x;
    const bool skip = meta && meta->wr1 > meta->wr0 && (r < meta->wr0 || r >= meta->wr1);
    const int wc0 = (meta && meta->wc1 > meta->wc0) ? meta->wc0 : 0, wc1 = (meta && meta->wc1 > meta->wc0) ? meta->wc1 : d.D;
    for (int c = threadIdx.x; c < d.D; c += blockDim.x)
        out[(size_t)r * d.D + c] = (skip || c < wc0 || c >= wc1) ? 0.0 : S[(size_t)r * d.ldS + c];
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

// Sample-cell emission (likelihood batches, pkb_solve_batch): the same per-cell rules as k_copy_domain /
// k_emit_dense / k_emit_population, evaluated only at the K (row, col) cells the caller reads -- the
// dense [D][D] day is never materialised.  grid = ceil(K / 256), block = 256
__global__ void k_copy_domain_cells(const double* __restrict__ S, ChainDims d, const int* __restrict__ cells, int K, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) out[k] = S[(size_t)cells[2 * k] * d.ldS + cells[2 * k + 1]];
}
__global__ void k_emit_dense_cells(const double* __restrict__ S, ChainDims d, const StepMeta* __restrict__ meta, double negval, int prob_model,
                                   int strict, const int* __restrict__ cells, int K, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const double add = prob_model ? meta->add : 0.0;
    const int r = cells[2 * k], cc = cells[2 * k + 1];
    if ((meta->wr1 > meta->wr0 && (r < meta->wr0 || r >= meta->wr1)) || (meta->wc1 > meta->wc0 && (cc < meta->wc0 || cc >= meta->wc1))) {
        out[k] = 0.0;                                   // outside the rows / columns this step computed (all below PKB_SPEC_TAU)
        return;
    }
    const double v = S[(size_t)r * d.ldS + cc];
    const bool keep = strict ? (v > negval) : (v != 0.0 && !(v < negval));
    out[k] = keep ? v + add : 0.0;
}
__global__ void k_emit_population_cells(CohortArgs ca, ChainDims d, double r_number, double centre_extra, int add_centre, double negval,
                                        int first_day, const int* __restrict__ cells, int K, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int r = cells[2 * k], c = cells[2 * k + 1], mid = d.D / 2;
    double v;
    if (first_day) {
        const double p = ca.S[0][(size_t)r * d.ldS + c];
        const double t = (p != 0.0 && !(p < negval)) ? p : 0.0;
        v = (t * r_number) * ca.w[0];
    } else {
        double acc = 0.0;
        for (int j = 0; j < ca.n; ++j) acc += ca.S[j][(size_t)r * d.ldS + c] * ca.w[j];
        v = acc * r_number;
        v = (v != 0.0 && !(v < negval)) ? v : 0.0;
    }
    if (add_centre && r == mid && c == mid) v += centre_extra;
    out[k] = v;
}

// out[day][k] = G[day][cells[k]]   grid = ndays, block = 256
__global__ void k_sample(const double* __restrict__ G, int D, const int* __restrict__ cells, int K, double* __restrict__ out) {
    const double* g = G + (size_t)blockIdx.x * D * D;
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        out[(size_t)blockIdx.x * K + k] = g[(size_t)cells[2 * k] * D + cells[2 * k + 1]];
}

// Single-vector transform through the shared-memory FFT (diagnostics).
// grid = 1, block = T, dyn smem = fft_smem_bytes(plan)
__global__ void PKB_ROWS_LB k_fft_test(const cplx* __restrict__ in, cplx* __restrict__ out, int inverse, FftPlan plan) {
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    cplx* tws = x + plan.N;
    const int tid = threadIdx.x, T = blockDim.x, N = plan.N;
    fft_load_twiddles(tws, plan, tid, T);
    __syncthreads();
    if (!inverse) {
        fft_forward_from(x, tws, plan, tid, T, [&](int i) -> cplx { return in[i]; });
        for (int k = tid; k < N; k += T) out[k] = x[__ldg(&plan.perm[k])];
    } else {
        for (int k = tid; k < N; k += T) x[__ldg(&plan.perm[k])] = in[k];
        __syncthreads();
        fft_inverse_to(x, tws, plan, tid, T, [&](int i, cplx v) { out[i] = v; });
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
// grid = ndays*D, block = 256.  Each thread owns a contiguous segment of the row: count its
// non-zeros, one block-wide exclusive scan of the 256 counts, ordered write.
// meta (optional): only the columns [wc0, wc1) this day's step computed can hold a non-zero (k_emit_dense, sparse_only)
__global__ void k_coo_write(const double* __restrict__ G, int D, const long long* __restrict__ rowoff, const int* __restrict__ rownnz,
                            int* __restrict__ rows, int* __restrict__ cols, double* __restrict__ vals, const StepMeta* __restrict__ meta) {
    PKB_SHARED(int, cnt, 256);
    const int r = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    if (rownnz[r] == 0) return;        // thresholded solutions are mostly empty rows: do not even read them
    const int w0 = (meta && meta->wc1 > meta->wc0) ? meta->wc0 : 0, w1 = (meta && meta->wc1 > meta->wc0) ? meta->wc1 : D;
    const int seg = (w1 - w0 + T - 1) / T;
    const int c0 = w0 + tid * seg, c1 = c0 + seg < w1 ? c0 + seg : w1;
    const double* row = G + (size_t)r * D;
    int n = 0;
    for (int c = c0; c < c1; ++c) n += row[c] != 0.0 ? 1 : 0;
    cnt[tid] = n;
    __syncthreads();
    // inclusive scan (Hillis-Steele) over the block
    for (int s = 1; s < T; s <<= 1) {
        const int add = tid >= s ? cnt[tid - s] : 0;
        __syncthreads();
        cnt[tid] += add;
        __syncthreads();
    }
    long long pos = rowoff[r] + cnt[tid] - n;
    const int rr = r % D;
    for (int c = c0; c < c1; ++c) {
        const double v = row[c];
        if (v != 0.0) {
            if (rows) rows[pos] = rr;       // (CSR output carries the row offsets instead)
            cols[pos] = c; vals[pos] = v;
            ++pos;
        }
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
    double pmax = -INFINITY, ks = 0.0, pab = 0.0, rab = 0.0;
    int kc = 0;
    for (int c = threadIdx.x; c < d.P; c += blockDim.x) {
        const double v = S[(size_t)r * d.ldS + c];
        rab = fmax(rab, fabs(v));
        if (r >= d.D || c >= d.D) { pmax = fmax(pmax, v); pab = fmax(pab, fabs(v)); }
        else if (!(v < negval)) { ks += v; kc += 1; }
    }
    const double brab = block_max(rab, red);
    const double bpmax = block_max(pmax, red);
    const double bks = block_sum(ks, red);
    const double bkc = block_sum((double)kc, red);
    const double bab = block_max(pab, red);
    if (threadIdx.x == 0) {
        RowStats rs;
        rs.padmax = bpmax; rs.ksum = bks; rs.padabs = bab; rs.kcnt = (int)bkc; rs.has_e = brab >= PKB_SPEC_TAU ? 1 : 0;
        rstat[r] = rs;
    }
}

}  // namespace pkb
