// One forward solve spread over the G GPUs of a box (BASELINE config 4 "single
// solve vs 1/2/4/8 GPUs"): a slab-decomposed version of the spectral-resident chain step of chain.cuh.
//
// The reference's chain state lives in Fourier space and is multiplied by each day's kernel spectrum
// (CalcSol.py:66,189-201).  While nothing of consequence lies outside the domain (the criterion of chain.cuh,
// PKB_SPEC_EPS, checked every day on the global maximum) the fold mod P moves nothing of consequence either, and one
// day is
//     columns:  S_hat[:, c] *= FFT_col(K_rowspectra[:, c]);  W[:, c] = IFFT_col(S_hat[:, c])     for every spectral column c
//     rows:     s[r, :]     = IFFT_row(W[r, :])                                                  for every needed row r
// The columns are independent of each other, and so are the rows: rank g owns the spectral columns [c0, c1) -- its
// slice of S_hat never leaves it -- and the linear-convolution rows [j0, j1).  Between the two passes every rank
// hands every other rank the (own columns) x (their rows) block of W: ONE all-to-all per day (NCCL over NVLink,
// 8 N (P + 2m) / G bytes out of every GPU), laid out by k_cols_dist so that each destination's block is contiguous.
// The per-day sums for the renormalisation (CalcSol.py:134-135) and the outside-domain maximum are combined with one
// all-gather of four doubles per rank.  The real state is only ever produced for output: no fold, no forward pass.
#pragma once
#include "chain.cuh"

namespace pkb {

struct DistGeom {
    int G, rank;
    int Cg;          // spectral columns per rank (a multiple of PKB_CB; the last rank's share may be short or empty)
    int c0, c1;      // own columns
    int mmax;        // the row map keeps linear rows [0, P + mmax) and [N - mmax, N): J = P + 2 mmax rows
    int J, Jr;       // needed rows, rows per rank (even)
    int j0, j1;      // own rows [j0, j1) of the J needed ones
    unsigned jr_magic;
    unsigned long long blk;   // complex elements of one (source, destination) block = spec_size(Cg, Jr)
};

// needed-row index of linear row i (or -1), and back
__host__ __device__ __forceinline__ int dist_row_to_j(const DistGeom& g, int i, int P, int N) {
    if (i < P + g.mmax) return i;
    if (i >= N - g.mmax) return i - (N - g.mmax) + (P + g.mmax);
    return -1;
}
__host__ __device__ __forceinline__ int dist_j_to_row(const DistGeom& g, int j, int P, int N) {
    return j < P + g.mmax ? j : j - (P + g.mmax) + (N - g.mmax);
}

// k_cols for the rank's own columns.  from_yt != 0: the state column is transformed from its row spectra Yt (first
// day: the state is the placed kernel of day 0); else it is read from Shat.  The product spectrum always goes back to
// Shat, the column-transformed product to `send`: block h (destination rank) holds rows [h Jr, (h + 1) Jr) of the own
// columns in the tiled layout of chain.cuh with ld = Jr.
__global__ void PKB_COLS_LB k_cols_dist(const cplx* __restrict__ Yt, const cplx* __restrict__ Krt, int m, ChainDims d, cplx* __restrict__ send,
                                        cplx* __restrict__ scr, FftPlan plan, cplx* __restrict__ Shat, size_t hstride, DistGeom g, int from_yt) {
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    cplx* tws = x + plan.N;
    const int tid = threadIdx.x, T = blockDim.x;
    fft_load_twiddles(tws, plan, tid, T);
    __syncthreads();
    const int lim = d.P;
    const int N = d.N, nq = 2 * m + 1;
    const int L = plan.nstage, R0 = plan_radix(plan, 0), RL = plan_radix(plan, L - 1), nbl = N / RL;
    cplx* myscr = scr + (size_t)blockIdx.x * ((size_t)plan.cols_kb * RL * T);
    const int off_last = plan.ntw - 1;
    for (int c = g.c0 + blockIdx.x; c < g.c1; c += gridDim.x) {
        const cplx* ycol = Yt + spec_index(c, 0, d.ldY);
        const cplx* kcol = Krt + spec_index(c, 0, d.ldK);
        cplx* hcol = Shat + (size_t)(c - g.c0) * hstride;
        const int cl = c - g.c0;
        auto ld_filter = [&](int i) -> cplx {
            if (i <= m) return kcol[(size_t)i * PKB_CB];
            if (i >= N - m) return kcol[(size_t)(i - (N - nq)) * PKB_CB];
            return cmake(0.0, 0.0);
        };
        auto ld_state = [&](int i) -> cplx { return i < lim ? ycol[(size_t)i * PKB_CB] : cmake(0.0, 0.0); };
        auto st_out = [&](int i, cplx v) {
            const int j = dist_row_to_j(g, i, d.P, N);
            if (j < 0) return;
            const int h = fast_div(j, g.Jr, g.jr_magic);
            send[(size_t)h * g.blk + spec_index(cl, j - h * g.Jr, g.Jr)] = v;
        };
        for (int phase = 0; phase < 2; ++phase) {
            auto ld = [&](int i) -> cplx { return phase ? ld_state(i) : ld_filter(i); };
            if (phase && !from_yt) {
                // the state column's spectrum is in Shat
            } else if (L == 1) {
                for (int i = tid; i < N; i += T) x[i] = ld(i);
                __syncthreads();
            } else {
                fft_stage_dispatch<false, true>(R0, N, N, plan.tw0, tid, T, ld, SmemStore{x});
                __syncthreads();
                fft_fwd_stages(x, tws, plan, 1, L - 1, N / R0, 0, tid, T);
            }
            const int fmode = !phase ? 0 : (from_yt ? 3 : 2);
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

// Per-rank, per-day sums: [0] kept sum, [1] kept count, [2] max outside the domain (signed, the boundary flag
// compares it with 1e-8), [3] max |value| outside the domain (the spectral criterion)
#define PKB_DIST_NSTAT 4

// Inverse row pass of the rank's own rows.  recv: block s (source rank) holds the own rows of rank s's columns.
// One job = one local row pair (2q, 2q + 1) carried as real / imaginary part of one transform.  Sloc: [Jr][ldL]
// real rows (domain columns only); rstat: [Jr] per-row statistics; the CTA that finishes last reduces them into
// stats[PKB_DIST_NSTAT].  No fold: what would fold is below PKB_SPEC_EPS by the criterion.
__global__ void PKB_ROWS_LB k_rows_inv_dist(const cplx* __restrict__ recv, ChainDims d, double* __restrict__ Sloc, int ldL, RowStats* __restrict__ rstat,
                                            double negval, FftPlan plan, int* __restrict__ done, double* __restrict__ stats, DistGeom g) {
    PKB_DYN_SMEM(raw);
    cplx* x = reinterpret_cast<cplx*>(raw);
    cplx* tws = x + plan.N;
    double* red = reinterpret_cast<double*>(raw);
    int* last = reinterpret_cast<int*>(red + PKB_RED_DOUBLES);
    const int tid = threadIdx.x, T = blockDim.x;
    fft_load_twiddles(tws, plan, tid, T);
    __syncthreads();
    const int P = d.P, N = d.N, D = d.D, Nc = d.Nc;
    const int nloc = g.j1 - g.j0;
    const int njobs = (nloc + 1) / 2;
    const double scale = 1.0 / ((double)N * (double)N);
    const cplx zero = cmake(0.0, 0.0);
    const unsigned cg_magic = div_magic(g.Cg);
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        const int ra = 2 * job, rb = ra + 1 < nloc ? ra + 1 : -1;
        const int npair = (Nc + 1) / 2;
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
                    const int c = 2 * col[u];
                    const int s = fast_div(c, g.Cg, cg_magic);
                    const cplx* base = recv + (size_t)s * g.blk + spec_index(c - s * g.Cg, 0, g.Jr);
                    ld_pair(base + (size_t)ra * PKB_CB, a[2 * u], a[2 * u + 1]);
                    if (rb >= 0) ld_pair(base + (size_t)rb * PKB_CB, b[2 * u], b[2 * u + 1]);
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
                            x[pk] = cmake(av.x, bv.x);
                        } else {
                            x[pk] = cmake(av.x - bv.y, av.y + bv.x);
                            x[pn] = cmake(av.x + bv.y, bv.x - av.y);
                        }
                    }
                }
            }
        }
        __syncthreads();
        fft_inverse_to(x, tws, plan, tid, T, SmemStore{x});
        __syncthreads();
        // rows of the domain: linear row i = j < D; everything else (pad rows, wrapped rows) only feeds the statistics
        const int ia = dist_j_to_row(g, g.j0 + ra, P, N), ib = rb >= 0 ? dist_j_to_row(g, g.j0 + rb, P, N) : -1;
        const bool dom_a = ia < D, dom_b = ib >= 0 && ib < D;
        double st[8] = {-INFINITY, 0.0, 0.0, 0.0, -INFINITY, 0.0, 0.0, 0.0};
        double* dst_a = Sloc + (size_t)ra * ldL;
        double* dst_b = Sloc + (size_t)(rb >= 0 ? rb : ra) * ldL;
        for (int c = tid; c < N; c += T) {
            const cplx z = x[c];
            const double va = z.x * scale;
            if (c < D) dst_a[c] = va;
            if (!dom_a || c >= D) { st[0] = fmax(st[0], va); st[3] = fmax(st[3], fabs(va)); }
            else if (!(va < negval)) { st[1] += va; st[2] += 1.0; }
            if (rb >= 0) {
                const double vb = z.y * scale;
                if (c < D) dst_b[c] = vb;
                if (!dom_b || c >= D) { st[4] = fmax(st[4], vb); st[7] = fmax(st[7], fabs(vb)); }
                else if (!(vb < negval)) { st[5] += vb; st[6] += 1.0; }
            }
        }
        __syncthreads();
        const double r = block_reduce8(st, red, tid, T);
        if (tid < 8) red[64 + tid] = r;
        __syncthreads();
        if (tid == 0 || (tid == 1 && rb >= 0)) {
            const double* q = red + 64 + 4 * tid;
            RowStats rs;
            rs.padmax = q[0]; rs.ksum = q[1]; rs.kcnt = (int)q[2]; rs.padabs = q[3]; rs.has_e = 0;
            rstat[tid ? rb : ra] = rs;
        }
        __syncthreads();
    }
    if (tid == 0) {
        __threadfence();
        last[0] = atomicAdd(done, 1) == (int)gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (last[0]) {
        __threadfence();
        // fixed order: contiguous chunks of rows per thread, then the block reduction
        double st[8] = {-INFINITY, 0.0, 0.0, 0.0, -INFINITY, 0.0, 0.0, 0.0};
        const int chunk = (nloc + T - 1) / T;
        const volatile RowStats* rv = rstat;
        for (int r = tid * chunk; r < nloc && r < (tid + 1) * chunk; ++r) {
            st[0] = fmax(st[0], rv[r].padmax);
            st[3] = fmax(st[3], rv[r].padabs);
            st[1] += rv[r].ksum;
            st[2] += (double)rv[r].kcnt;
        }
        __syncthreads();
        const double rr = block_reduce8(st, red, tid, T);
        if (tid < 4) {
            // [ksum, kcnt, padmax, padabs]
            const int slot = tid == 0 ? 2 : (tid == 1 ? 0 : (tid == 2 ? 1 : 3));
            stats[slot] = (nloc > 0 || (tid != 0)) ? rr : -INFINITY;
        }
        if (tid == 0) *done = 0;
    }
}

// r_small_vals(prob_model=True) of the rank's domain rows with the GLOBAL sums (CalcSol.py:112-136): allstats holds
// the PKB_DIST_NSTAT values of every rank in rank order (summed in that order: identical on every rank).
// out: [Jr][D] rows of this day (rows beyond the domain stay zero); meta (rank-independent): [ksum, kcnt, padmax, padabs].
// grid = Jr, block = 256
__global__ void k_emit_dist(const double* __restrict__ Sloc, int ldL, ChainDims d, const double* __restrict__ allstats, DistGeom g, double negval,
                            double* __restrict__ out, double* __restrict__ meta, double* __restrict__ worst) {
    double ks = 0.0, kc = 0.0, pm = -INFINITY, pa = 0.0;
    for (int r = 0; r < g.G; ++r) {
        const double* s = allstats + (size_t)r * PKB_DIST_NSTAT;
        ks += s[0]; kc += s[1]; pm = fmax(pm, s[2]); pa = fmax(pa, s[3]);
    }
    const double add = (1.0 - ks) / kc;
    const int jl = blockIdx.x;
    if (jl == 0 && threadIdx.x == 0) {
        if (meta) { meta[0] = ks; meta[1] = kc; meta[2] = pm; meta[3] = pa; }
        if (worst) worst[0] = fmax(worst[0], pa);           // (one writer per rank and day, stream-ordered)
    }
    const int i = g.j0 + jl;                                 // j = i for domain rows
    double* dst = out + (size_t)jl * d.D;
    if (jl >= g.j1 - g.j0 || i >= d.D) {
        for (int c = threadIdx.x; c < d.D; c += blockDim.x) dst[c] = 0.0;
        return;
    }
    const double* src = Sloc + (size_t)jl * ldL;
    for (int c = threadIdx.x; c < d.D; c += blockDim.x) {
        const double v = src[c];
        dst[c] = (v != 0.0 && !(v < negval)) ? v + add : 0.0;
    }
}

}  // namespace pkb
