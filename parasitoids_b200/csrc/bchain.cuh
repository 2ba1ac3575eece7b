// Batched chain kernels of the likelihood batch (pkb_solve_batch): ONE launch per pass and step for a whole
// group of proposals.
//
// A Kalbar-sized chain step (801^2 domain, torus ~1000-1500) is three short persistent kernels of one or two
// rounds each (chain.cuh); launched per proposal their tails and launch gaps cost more than the transforms.  The
// proposals of a likelihood batch are independent solves of the SAME day count (Bayes_Run.py:204-336 evaluates
// one proposal per call), so here step n of every proposal of a group is enqueued together: each kernel below
// walks a job table  [proposal 0 jobs | proposal 1 jobs | ...]  with one persistent grid, every proposal with its
// own geometry (torus, support window, kernel radius, FFT plan) in a device-resident descriptor (BStep).  The job
// bodies are those of k_kernel_rows / k_rows_fwd / k_cols / k_rows_inv, job for job and row for row, including the
// device-side switch to the truncated-source torus (TruncGeom) -- so a proposal's result has the same bits as on
// the per-proposal path whenever that one takes exact steps (no spectral-resident steps, no tau windows).
//
//   kb_rows_fwd   per proposal: m + 1 kernel-row jobs (row spectra of the day's kernel), then the forward row
//                 jobs of the state
//   kb_cols       per proposal: one job per spectral column (filter column FFT, state column FFT, product, inverse)
//   kb_rows_inv   per proposal: inverse row jobs, fold mod P, new state, per-row statistics
//   kb_finish     one CTA per proposal: flag / sums of the step (step_finalize_block, same summation order as the
//                 finalising CTA of k_rows_inv at this CTA size), then the sample-cell emission of the day
//   kb_init       day 0: the first kernel placed at the domain centre, control block of a fresh state
#pragma once
#include "chain.cuh"

#define PKB_BT 128          // CTA size of every batched chain kernel (what get_plan picks whenever >= 4 transforms fit an SM)
#ifndef PKB_BCH_B
#define PKB_BCH_B 4         // resident CTAs per SM the batched kernels are compiled for (register cap 65536 / (PKB_BT * PKB_BCH_B))
#endif
#define PKB_BCH_LB __launch_bounds__(PKB_BT, PKB_BCH_B)

namespace pkb {

struct BStep {
    ChainDims d;            // geometry of this step: the proposal's P torus on the step's own FFT torus, or a support window
    TruncGeom tg;           // second geometry for a truncated source (tg.N == 0: none)
    int m, Wk;              // kernel radius, side of the kernel window K
    int plan, plan_t;       // indices into the plan table
    int cols_kb;            // k_cols geometry of `plan` at PKB_BT threads (tg.cols_kb for plan_t)
    int pad_;
    const double* K;        // the day's kernel (Wk x Wk window centred on the release cell)
    const double* src;      // state before the step
    double* dst;            // state after it
    cplx* Yt;               // half-spectrum rows of the state, tiled transposed
    cplx* Wt;               // product spectrum after the inverse column transforms
    cplx* Krt;              // row spectra of the kernel (on whichever torus the step runs)
    RowStats* rstat;
    ChainCtrl* ctrl;
    StepMeta* meta;         // the day's record
    int* hint;              // set by kb_rows_inv once a DOMAIN row of the new state shows a cell above the flag threshold outside the domain
};

// sample-cell emission of one (proposal, day): mode 0 copy (probability model day 0, Run.py:454-458), 1 r_small_vals +
// renormalisation (CalcSol.py:112-136), 2 population first day (CalcSol.py:236-237), 3 population later days
// (CalcSol.py:322-323, one cohort), -1 nothing (the dropped leading spread day)
struct BEmit {
    const double* S;
    const StepMeta* meta;
    double* out;            // [K]
    int ldS, mode;
    double w0, centre_extra;
};

struct BInit {
    const double* K;
    double* S;
    ChainCtrl* ctrl;
    int Wk, m, ldS, D;
};

struct BShared {
    BStep st;
    FftPlan plan;
    int tr, c_trunc, skip;
};

// Descriptor of proposal p -> shared memory, geometry switched to the truncated-source torus if the source state is
// truncated (the decision every kernel of the step takes from the same control block), plan -> shared memory, base
// twiddles loaded.  Returns synchronised.
__device__ __forceinline__ void b_select(BShared* sh, cplx* x, const BStep* __restrict__ steps, const FftPlan* __restrict__ plans, int p, int tid,
                                         int T) {
    __syncthreads();                                    // nobody still reads the previous descriptor / twiddles
    {
        const int* s = reinterpret_cast<const int*>(steps + p);
        int* o = reinterpret_cast<int*>(&sh->st);
        for (int i = tid; i < (int)(sizeof(BStep) / sizeof(int)); i += T) o[i] = s[i];
    }
    __syncthreads();
    if (tid == 0) {
        const int c_trunc = reinterpret_cast<const volatile ChainCtrl*>(sh->st.ctrl)->trunc;
        const int tr = (sh->st.tg.N && !sh->st.d.win && c_trunc) ? 1 : 0;
        sh->c_trunc = c_trunc;
        sh->tr = tr;
        if (tr) {
            sh->st.d.N = sh->st.tg.N; sh->st.d.Nc = sh->st.tg.Nc; sh->st.d.ldW = sh->st.tg.ldW;
            sh->st.plan = sh->st.plan_t;
            sh->st.cols_kb = sh->st.tg.cols_kb;
        }
    }
    __syncthreads();
    {
        const int* s = reinterpret_cast<const int*>(plans + sh->st.plan);
        int* o = reinterpret_cast<int*>(&sh->plan);
        for (int i = tid; i < (int)(sizeof(FftPlan) / sizeof(int)); i += T) o[i] = s[i];
    }
    __syncthreads();
    fft_load_twiddles(x + sh->plan.N, sh->plan, tid, T);
    __syncthreads();
}

// grid = persistent over job0[nprob] jobs, block = PKB_BT.  Jobs of proposal p: [job0[p], job0[p] + m + 1) kernel rows,
// then (rows + 1) / 2 forward row pairs (host upper bound; a truncated source has fewer).
__global__ void PKB_BCH_LB kb_rows_fwd(const BStep* __restrict__ steps, const FftPlan* __restrict__ plans, const int* __restrict__ job0, int nprob) {
    PKB_DYN_SMEM(raw);
    PKB_SHARED(BShared, shm, 1);
    BShared* sh = shm;
    cplx* x = reinterpret_cast<cplx*>(raw);
    const int tid = threadIdx.x, T = blockDim.x;
    const int total = job0[nprob];
    int p = -1, pn = 0;
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
        while (g >= job0[pn + 1]) ++pn;             // jobs of a CTA increase monotonically
        if (pn != p) {
            p = pn;
            b_select(sh, x, steps, plans, p, tid, T);
        }
        const FftPlan& plan = sh->plan;
        const ChainDims& d = sh->st.d;
        cplx* tws = x + plan.N;
        const int m = sh->st.m;
        int job = g - job0[p];
        if (job <= m) {
            kernel_rows_job(x, tws, sh->st.K, sh->st.Wk, m, d, sh->st.Krt, plan, job, tid, T);
            continue;
        }
        job -= m + 1;
        const int lim = d.win ? d.wn : (sh->c_trunc ? d.D : d.P);
        if (2 * job >= lim) continue;
        const double* S = sh->st.src;
        if (d.win) S += (size_t)d.wr0 * d.ldS + d.wc0;     // rows / columns below are relative to the window
        const int r0 = 2 * job;
        const bool two = r0 + 1 < lim;
        const double* s0 = S + (size_t)r0 * d.ldS;
        const double* s1 = s0 + (two ? d.ldS : 0);
        auto ld = [&](int j) -> cplx {
            if (j >= lim) return cmake(0.0, 0.0);
            return cmake(s0[j], two ? s1[j] : 0.0);
        };
        fft_forward_from(x, tws, plan, tid, T, ld, false);
        unpack_store(x, plan, d.Nc, sh->st.Yt, d.ldY, r0, two, tid, T);
        __syncthreads();
    }
}

// grid = persistent, block = PKB_BT.  Jobs of proposal p: its spectral columns (Nc of the step's own torus; the
// truncated-source torus has fewer).  scr: gridDim.x slices of scr_per_cta complex.
__global__ void PKB_BCH_LB kb_cols(const BStep* __restrict__ steps, const FftPlan* __restrict__ plans, const int* __restrict__ job0, int nprob,
                                    cplx* __restrict__ scr, size_t scr_per_cta) {
    PKB_DYN_SMEM(raw);
    PKB_SHARED(BShared, shm, 1);
    BShared* sh = shm;
    cplx* x = reinterpret_cast<cplx*>(raw);
    const int tid = threadIdx.x, T = blockDim.x;
    const int total = job0[nprob];
    cplx* myscr = scr + (size_t)blockIdx.x * scr_per_cta;
    int p = -1, pn = 0;
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
        while (g >= job0[pn + 1]) ++pn;
        if (pn != p) {
            p = pn;
            b_select(sh, x, steps, plans, p, tid, T);
        }
        const FftPlan& plan = sh->plan;
        const ChainDims& d = sh->st.d;
        const int c = g - job0[p];
        if (c >= d.Nc) continue;
        cplx* tws = x + plan.N;
        const int m = sh->st.m;
        const bool tr = sh->tr != 0;
        const int lim = d.win ? d.wn : (sh->c_trunc ? d.D : d.P);
        const int N = d.N, nq = 2 * m + 1;
        const int L = plan.nstage, R0 = plan_radix(plan, 0), RL = plan_radix(plan, L - 1), nbl = N / RL;
        const int hi = (d.win ? d.wn : (tr ? d.D : d.P)) + m;   // rows [0, extent + m) and [N-m, N) are needed downstream
        const int off_last = plan.ntw - 1;
        // row i of this column lives at col[i * PKB_CB]
        const cplx* ycol = sh->st.Yt + spec_index(c, 0, d.ldY);
        const cplx* kcol = sh->st.Krt + spec_index(c, 0, d.ldK);
        cplx* wcol = sh->st.Wt + spec_index(c, 0, d.ldW);
        auto ld_filter = [&](int i) -> cplx {
            if (i <= m) return kcol[(size_t)i * PKB_CB];
            if (i >= N - m) return kcol[(size_t)(i - (N - nq)) * PKB_CB];
            return cmake(0.0, 0.0);
        };
        auto ld_state = [&](int i) -> cplx { return i < lim ? ycol[(size_t)i * PKB_CB] : cmake(0.0, 0.0); };
        auto st_out = [&](int i, cplx v) {
            if (i < hi || i >= N - m) wcol[(size_t)i * PKB_CB] = v;
        };
        for (int phase = 0; phase < 2; ++phase) {
            auto ld = [&](int i) -> cplx { return phase ? ld_state(i) : ld_filter(i); };
            if (L == 1) {
                for (int i = tid; i < N; i += T) x[i] = ld(i);
                __syncthreads();
            } else {
                fft_stage_dispatch<false, true>(R0, N, N, plan.tw0, tid, T, ld, SmemStore{x});
                __syncthreads();
                fft_fwd_stages(x, tws, plan, 1, L - 1, N / R0, 0, tid, T);
            }
#define PKB_CALL_(RR) cols_final<RR>(x, myscr, (cplx*)nullptr, tid, T, nbl, phase)
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

// grid = persistent, block = PKB_BT.  The job table has TWO entries per proposal: [0, nprob) the inverse row jobs of
// k_rows_inv (host upper bound over both geometries), [nprob, 2 nprob) -- truncated-source torus only -- the jobs whose
// output rows all lie outside the domain, which that geometry takes last (rows ascending).  Once a domain row has shown a pad
// cell above the flag threshold the step is flagged whatever those rows hold (CalcSol.py:36-37), and a flagged state is read
// as truncated to the domain (:200-201): nobody will read them.  The first pass sets BStep::hint, the second one -- a whole
// pass over the group later -- skips its jobs when it is set: 38 % of the inverse row jobs of a Kalbar-sized flagged step.
// The step's flag / sums are reduced by kb_finish.
__global__ void PKB_BCH_LB kb_rows_inv(const BStep* __restrict__ steps, const FftPlan* __restrict__ plans, const int* __restrict__ job0, int nprob,
                                        double negval) {
    PKB_DYN_SMEM(raw);
    PKB_SHARED(BShared, shm, 1);
    BShared* sh = shm;
    cplx* x = reinterpret_cast<cplx*>(raw);
    double* red = reinterpret_cast<double*>(raw);       // reduction scratch: start of the transform buffer, idle when used
    const int tid = threadIdx.x, T = blockDim.x;
    const int total = job0[2 * nprob];
    const cplx zero = cmake(0.0, 0.0);
    int p = -1, pn = 0;
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
        while (g >= job0[pn + 1]) ++pn;
        if (pn != p) {
            p = pn;
            b_select(sh, x, steps, plans, p < nprob ? p : p - nprob, tid, T);
            if (p >= nprob) {
                if (tid == 0) sh->skip = *reinterpret_cast<volatile int*>(sh->st.hint);
                __syncthreads();
            }
        }
        const int part = p < nprob ? 0 : 1;
        const FftPlan& plan = sh->plan;
        const ChainDims& d = sh->st.d;
        cplx* tws = x + plan.N;
        const int m = sh->st.m;
        const bool tr = sh->tr != 0;
        const int P = d.P, N = d.N, D = d.D, Nc = d.Nc;
        const int wout = d.wn + 2 * m;                       // window mode: side of the result
        const int njobs = d.win ? (wout + 1) / 2 : (tr ? rows_inv_jobs_trunc(P, D, m) : rows_inv_jobs(P, m));
        const int ndom = tr ? (D + 1) / 2 : njobs;          // truncated-source torus: the row pairs that start inside the domain come first
        int job = g - job0[p];
        if (part == 0 && job >= ndom) continue;
        if (part == 1) job += ndom;
        if (job >= njobs) continue;
        const double scale = 1.0 / ((double)N * (double)N);
        const int E = D + m, Lo = P - m;                     // truncated mode: extent of the positive rows / columns, first folded one
        const cplx* Wt = sh->st.Wt;
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
        } else {
            rows_inv_decode(job, m, P, N, ra, rb, out_a, out_b, fold);
        }
        if (part == 1 && sh->skip) {
            // (see the kernel's header: the step is flagged whatever these rows hold; only its recorded pad maximum differs)
            if (tid == 0 || (tid == 1 && out_b >= 0)) {
                RowStats rs;
                rs.padmax = 1.0; rs.ksum = 0.0; rs.padabs = 1.0; rs.kcnt = 0; rs.has_e = 0;
                sh->st.rstat[tid ? out_b : out_a] = rs;
            }
            continue;
        }
        // scatter the Hermitian pair Z = A + iB into digit-reversed order; each thread
        // handles two adjacent columns (32-byte loads per row)
        const int npair = zjob ? 0 : (Nc + 1) / 2;
        for (int j0 = tid; j0 < npair; j0 += PKB_UNPACK_U * T) {
            int4 pr[PKB_UNPACK_U];
            int col[PKB_UNPACK_U];
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
                        const int pk = h ? pr[u].z : pr[u].x, pn2 = h ? pr[u].w : pr[u].y;
                        if (k == 0 || N - k == k) {
                            x[pk] = cmake(av.x, bv.x);        // self-conjugate bins are real
                        } else {
                            x[pk] = cmake(av.x - bv.y, av.y + bv.x);    // A + iB
                            x[pn2] = cmake(av.x + bv.y, bv.x - av.y);   // conj(A) + i conj(B)
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
        const int ncols = d.win ? wout : P;
        const int col0 = d.win ? d.wc0 - m : 0;
        double* Sout = sh->st.dst;
        double* dst_a = Sout + (size_t)out_a * d.ldS + col0;
        double* dst_b = Sout + (size_t)(out_b >= 0 ? out_b : out_a) * d.ldS + col0;
        // st[0..3]: row a (pad max, kept sum, kept count, -min); st[4..7]: row b
        double st[8] = {-INFINITY, 0.0, 0.0, 0.0, -INFINITY, 0.0, 0.0, 0.0};
        if (d.win) { st[0] = st[4] = 0.0; }                          // the untouched rest of the row is zero
        const bool pad_a = out_a >= D, pad_b = out_b >= D;
        const int Dc = D - col0;                                     // first pad column, relative to col0
        bool e_a = false, e_b = false;                               // a cell >= PKB_SPEC_TAU in this row (RowStats::has_e)
        unsigned cmask = 0;                                          // (column extents are not tracked here: exact support windows)
        if (tr) rows_inv_emit<2>(x, dst_a, dst_b, out_b >= 0, fold, zjob, ncols, m, N, P, E, Lo, pad_a, pad_b, Dc, scale, negval, st, e_a, e_b, cmask, (int*)nullptr, tid, T);
        else if (d.win) rows_inv_emit<1>(x, dst_a, dst_b, out_b >= 0, fold, zjob, ncols, m, N, P, E, Lo, pad_a, pad_b, Dc, scale, negval, st, e_a, e_b, cmask, (int*)nullptr, tid, T);
        else rows_inv_emit<0>(x, dst_a, dst_b, out_b >= 0, fold, zjob, ncols, m, N, P, E, Lo, pad_a, pad_b, Dc, scale, negval, st, e_a, e_b, cmask, (int*)nullptr, tid, T);
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
            sh->st.rstat[tid ? out_b : out_a] = rs;
            if (!d.win && (tid ? out_b : out_a) < D && rs.padmax > 1e-8) *reinterpret_cast<volatile int*>(sh->st.hint) = 1;
        }
        __syncthreads();
    }
}

// value of one sample cell under the emission rule of `e` (the per-cell rules of k_copy_domain_cells /
// k_emit_dense_cells / k_emit_population_cells)
__device__ __forceinline__ double b_emit_cell(const BEmit& e, int r, int c, int D, double r_number, double negval) {
    const double v0 = e.S[(size_t)r * e.ldS + c];
    if (e.mode == 0) return v0;
    if (e.mode == 1) {
        const StepMeta* meta = e.meta;
        if ((meta->wr1 > meta->wr0 && (r < meta->wr0 || r >= meta->wr1)) || (meta->wc1 > meta->wc0 && (c < meta->wc0 || c >= meta->wc1)))
            return 0.0;                                  // outside the rows / columns this step computed
        const bool keep = v0 != 0.0 && !(v0 < negval);
        return keep ? v0 + meta->add : 0.0;
    }
    const int mid = D / 2;
    double v;
    if (e.mode == 2) {
        const double t = (v0 != 0.0 && !(v0 < negval)) ? v0 : 0.0;
        v = (t * r_number) * e.w0;
        if (r == mid && c == mid) v += e.centre_extra;
    } else {
        double acc = 0.0;
        acc += v0 * e.w0;
        v = acc * r_number;
        v = (v != 0.0 && !(v < negval)) ? v : 0.0;
    }
    return v;
}

// grid = nprob, block = PKB_BT.  steps == nullptr: emission only (day 0).
__global__ void kb_finish(const BStep* __restrict__ steps, const BEmit* __restrict__ emits, const int* __restrict__ cells, int K, int D,
                          double r_number, double negval) {
    PKB_SHARED(double, red, PKB_RED_DOUBLES);
    const int p = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    if (steps) {
        const BStep& s = steps[p];
        SpecIn si = {0, 0, 0.0, 0.0, 0, 0, 0, 0, 0, nullptr, nullptr, 0};
        if (s.d.win) {
            const int wout = s.d.wn + 2 * s.m;
            si.rowwin = 1; si.w0 = s.d.wr0 - s.m; si.w1 = si.w0 + wout;
            si.c0 = s.d.wc0 - s.m; si.c1 = si.c0 + wout;
        }
        step_finalize_block(s.rstat, s.d, s.ctrl, s.meta, 1, red, tid, T, si, 1e-8);
        if (tid == 0) { s.ctrl->fused = 0; *s.hint = 0; }
        __syncthreads();                                 // (the emission below reads the record thread 0 just wrote)
    }
    const BEmit e = emits[p];
    if (e.mode < 0) return;
    for (int k = tid; k < K; k += T) e.out[k] = b_emit_cell(e, cells[2 * k], cells[2 * k + 1], D, r_number, negval);
}

// grid = (2 * mmax + 1, nprob), block = 128: first kernel of every proposal at the centre of its (zeroed) state
__global__ void kb_init(const BInit* __restrict__ inits) {
    const BInit b = inits[blockIdx.y];
    if ((int)blockIdx.x > 2 * b.m) return;
    const int ck = b.Wk / 2, cd = b.D / 2;
    const int dy = (int)blockIdx.x - b.m;
    for (int t = threadIdx.x; t < 2 * b.m + 1; t += blockDim.x) {
        const int dx = t - b.m;
        b.S[(size_t)(cd + dy) * b.ldS + cd + dx] = b.K[(size_t)(ck + dy) * b.Wk + ck + dx];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ChainCtrl* c = b.ctrl;
        c->trunc = 1; c->flag = 0; c->fused = 0;
        c->spec = 0; c->stored = 0; c->hint = 1; c->eps_sum = 0.0; c->eps_max = 0.0;
        c->er0 = c->er1 = c->ec0 = c->ec1 = 0;
    }
}

}  // namespace pkb
