// Phase 1 kernels: per-day dispersal kernel construction (ParasitoidModel.py).
//
//   k_bvn_setup      covariance constants + support half-width for mu = 0
//   k_hprob          h_flight_prob                   (ParasitoidModel.py:231-309)
//   k_drift          flight-averaged drift, cell offsets, per-period support
//                                                    (:439-495, :329-348)
//   k_period         BVN corner lattice -> cell masses -> clipped accumulation
//                                                    (:311-380, :497-558)
//   k_day_finalize   loss/total checks, local-diffusion blob, threshold +
//                    renormalise, crop radius        (:562-613, CalcSol.py:112-136)
//   k_mvn_cdf        stand-alone get_mvn_cdf_values  (:311-380)
//
// All arithmetic is IEEE fp64.  One CTA handles one take-off period of one day;
// the days of a solve (and the proposals of a batch) are the grid's y/z extent.
#pragma once
#include "bvn.cuh"
#if !PKB_IS_EMUL
#include <cooperative_groups.h>
#endif

namespace pkb {

#define PKB_ST_HPROB_RANGE 1      // hprob[t] outside [-1e-9, 1.000000001]  (:529)
#define PKB_ST_NEG_LOSS 2         // loss < 0                               (:569)
#define PKB_ST_PMF_NEG 4          // pmf.min() < -1e-8, first block         (:570)
#define PKB_ST_PMF_GT1 8          // pmfsum > 1.00001, first block          (:571)
#define PKB_ST_PMF_NEG2 16        // pmf.min() < -1e-8 after local blob     (:589)
#define PKB_ST_TOT_GT1 32         // total > 1.00001 after local blob       (:590)
#define PKB_ST_WARNED 64          // RuntimeWarning path taken              (:547-558)
#define PKB_ST_BORDERLINE 128     // ring-growth test within ring_tol of cdf_eps: decided by the reference-order running sum (:345-373)

#define PKB_CDF_EPS 0.001
#define PKB_LATTICE_CAP 5120      // doubles of shared memory for the corner lattice tile, at most (lattices up to CAP / 2 corners per side)
#define PKB_TILE_TARGET 3840      // ... and what k_period is launched with unless a lattice row pair needs more: with 6 nmax doubles of
                                  // marginals beside it, six 128-thread CTAs fit an SM (85 registers, __launch_bounds__(256, 3))
#define PKB_BVN_SEG 12            // lattice corners a thread marches along one column (k_period)

// Per-period contributions are accumulated EXACTLY and order-independently: every contribution x = h[t] * cdf (a
// double in [0, 1]) is split without error into hi = the multiple of 2^-50 nearest to x and the remainder
// lo = x - hi (|lo| <= 2^-51, exact in fp64), and the two parts are added to two 64-bit integer planes -- hi in
// units of 2^-50, lo in units of 2^-100 -- with integer atomics.  Integer addition is associative, so the day's
// window is bit-identical from run to run whatever order the period CTAs retire in (fp64 atomics gave sums that
// differed in the last bits, and with them the occasional keep/drop decision at the 1e-8 threshold), and nothing
// above 2^-100 = 8e-31 is ever lost: the sum is the correctly rounded one to within an ulp, closer to the exact
// value than the reference's own sequential `pmf += hprob * cdf` (ParasitoidModel.py:539).  Ranges: a cell's
// contributions sum to <= 1 (hi plane <= 2^50), and |lo| 2^100 <= 2^49 per period leaves room for 2^13 periods.
#define PKB_ACC_HI 1125899906842624.0                  // 2^50
#define PKB_ACC_LO 1267650600228229401496703205376.0   // 2^100
__device__ __forceinline__ void acc_add_exact(double* cell_hi, double* cell_lo, double v) {
    const double h = rint(v * PKB_ACC_HI);             // |v| <= 1: exact product, integer-valued double
    const double r = v - h * (1.0 / PKB_ACC_HI);       // exact (h / 2^50 is v rounded to a coarser grid)
    atomicAdd(reinterpret_cast<unsigned long long*>(cell_hi), (unsigned long long)(long long)h);
    const long long q = __double2ll_rn(r * PKB_ACC_LO);
    if (q) atomicAdd(reinterpret_cast<unsigned long long*>(cell_lo), (unsigned long long)q);
}
__device__ __forceinline__ double acc_exact_to_double(double bits_hi, double bits_lo) {
    return (double)__double_as_longlong(bits_hi) * (1.0 / PKB_ACC_HI) + (double)__double_as_longlong(bits_lo) * (1.0 / PKB_ACC_LO);
}

struct DayParams {      // one per (proposal, day) problem
    double lam, aw, bw, a1, b1, a2, b2;   // hparams (Run.py:377)
    double mu_r;
    double cell;                          // rad_dist / rad_res
    double sprd_factor, sdx, sdy;         // kind == 1: mixing weight and mean drift (metres) of the day-0 spread kernel
    int n_periods;
    int rad_res;
    int start_indx;                       // floor(start_time * periods), 0 if None
    int wind_day;                         // row of the wind array holding this day
    int has_next;                         // wind_day + 1 exists in the wind data
    int single;                           // 1-D wind row test form (:426-428)
    int bvn_S, bvn_Sl;                    // indices into the BvnPar array
    int kind;                             // 0: a day's dispersal kernel (prob_mass); 1: the local day-0 spread kernel of
                                          // Bayes_Run.py:245-270 / Bayes_MAP.py:247-277 (no wind, no threshold)
};

struct DayMeta {        // results of one (proposal, day) problem
    double loss, pmfsum, total, kept_sum, add;
    int rad;            // crop radius of the returned pmf
    int nnz;
    int status;
    int ext;            // max over periods of |cell offset| + support half-width
    int hl;             // support half-width of the local-diffusion blob
    int pad_;
};

// ---------------------------------------------------------------------------
// dpar[i] = (sig_x, sig_y, rho) if !is_cov, else (S00, S11, S01)
__global__ void k_bvn_setup(BvnPar* pars, const double* __restrict__ dpar /*[n][3]*/, const double* __restrict__ cell, int is_cov) {
    // one 32-thread block per covariance
    const int i = blockIdx.x;
    PKB_SHARED(int, found, 1);
    PKB_SHARED(int, okv, 32);
    BvnPar& p = pars[i];
    if (threadIdx.x == 0) {
        if (is_cov) bvn_setup_cov(p, dpar[3 * i], dpar[3 * i + 1], dpar[3 * i + 2]);
        else bvn_setup(p, dpar[3 * i], dpar[3 * i + 1], dpar[3 * i + 2]);
        found[0] = -1;
    }
    __syncthreads();
    const double c = cell[i];
    for (int base = 0; base < 4096; base += 32) {
        const int h = base + threadIdx.x;
        okv[threadIdx.x] = (1.0 - square_prob(p, c, h, 0.0, 0.0) < PKB_CDF_EPS) ? 1 : 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int t = 0; t < 32; ++t)
                if (okv[t]) { found[0] = base + t; break; }
        }
        __syncthreads();
        if (found[0] >= 0) break;
    }
    if (threadIdx.x == 0) p.h0 = found[0];   // -1 => not found below 4096 (caller reports)
}

// ---------------------------------------------------------------------------
// get_wind_data's interpolation (ParasitoidModel.py:162-227): every raw interval becomes interp_num periods,
// linearly blended with numpy's linspace weights w1 = i * (1 / interp_num), value = (1 - w1) * a + w1 * b, and
// windr = sqrt(wx^2 + wy^2) recomputed.  mode 0 ('00:00'): the day's last interval runs towards the next day's
// first sample; on the final day it repeats the last raw sample, raw windr included (:196-204).  mode 1
// ('00:30'): the day's first interval comes from the previous day's last sample; on the first day it repeats the
// first raw sample (:206-223).  raw: [nd][npts][3], out: [nd][npts * interp_num][3].  grid = (ceil(periods/256), nd)
__global__ void k_wind_interp(const double* __restrict__ raw, int nd, int npts, int interp_num, int mode, double* __restrict__ out) {
    const int day = blockIdx.y;
    const int periods = npts * interp_num;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= periods) return;
    const int k = t / interp_num, i = t - k * interp_num;       // output interval, position inside it
    const double step = 1.0 / (double)interp_num;
    const double w1 = (double)i * step, w0 = 1.0 - w1;
    const double* r = raw + (size_t)day * npts * 3;
    const double* a;
    const double* b;
    bool copy_a = false;          // the interval repeats sample a
    bool raw_r = false;           // ... including its raw windr
    if (mode == 0) {
        a = r + 3 * k;
        if (k + 1 < npts) b = a + 3;
        else if (day + 1 < nd) b = r + (size_t)npts * 3;
        else { b = a; copy_a = true; raw_r = true; }
    } else {
        if (k > 0) { a = r + 3 * (k - 1); b = r + 3 * k; }
        else if (day > 0) { a = r - 3; b = r; }
        else { a = r; b = r; copy_a = true; }
    }
    double x, y;
    if (copy_a) { x = a[0]; y = a[1]; }
    else { x = w0 * a[0] + w1 * b[0]; y = w0 * a[1] + w1 * b[1]; }
    double* o = out + ((size_t)day * periods + t) * 3;
    o[0] = x;
    o[1] = y;
    o[2] = raw_r ? a[2] : sqrt(x * x + y * y);
}

// ---------------------------------------------------------------------------
// h_flight_prob: grid = problems, block = 256, dyn smem = 4*periods doubles
// f_out / g_out (optional, one problem only): f_time_prob and g_wind_prob (:231-267)
__global__ void k_hprob(const DayParams* __restrict__ dps, const double* __restrict__ wind, int periods,
                        double* __restrict__ hprob_out, DayMeta* __restrict__ meta, double* __restrict__ f_out,
                        double* __restrict__ g_out) {
    PKB_DYN_SMEM(raw);
    PKB_SHARED(double, red, 256);
    double* f = reinterpret_cast<double*>(raw);   // likelihood -> f
    double* g = f + periods;
    double* c1 = g + periods;
    double* c2 = c1 + periods;
    const DayParams dp = dps[blockIdx.x];
    if (dp.kind) return;                          // (the spread kernel has no flight probability; hprob stays zero)
    const int n = dp.single ? 1 : periods;
    const double* w = wind + (size_t)dp.wind_day * periods * 3;
    const int tid = threadIdx.x, T = blockDim.x;
    // t_tild = linspace(0, 24 - 24/n, n)  (:260); numpy: start + i*step, last = stop
    const double stop = 24.0 - 24.0 / n;
    const double step = n > 1 ? stop / (n - 1) : 0.0;
    double part = 0.0, pmax = 0.0;
    for (int i = tid; i < n; i += T) {
        double t = (i == n - 1 && n > 1) ? stop : i * step;
        double up = 1.0 / (1.0 + exp(-dp.b1 * (t - dp.a1)));
        double dn = 1.0 / (1.0 + exp(-dp.b2 * (t - dp.a2)));
        double lik = fmax(up - dn, 0.0);
        f[i] = lik;
        part += lik;
        const double wr = dp.single ? w[2] : w[3 * i + 2];
        g[i] = 1.0 / (1.0 + exp(dp.bw * (wr - dp.aw)));   // g_wind_prob (:240)
    }
    const double tot = block_sum(part, red);
    for (int i = tid; i < n; i += T) {
        f[i] = f[i] / tot;                                // (:267)
        pmax = fmax(pmax, f[i]);
    }
    const double fmx = block_max(pmax, red);
    // two sequential prefix sums, same order as np.cumsum (:306-307)
    if (tid == 0) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) { acc += f[i]; c1[i] = acc; }
        acc = 0.0;
        for (int i = 0; i < n; ++i) {
            const double fg = f[i] * g[i];
            acc += (1.0 - c1[i]) * (f[i] - fg);
            c2[i] = acc;
        }
    }
    __syncthreads();
    int bad = 0;
    for (int i = tid; i < n; i += T) {
        const double fg = f[i] * g[i];
        const double tv = (double)(i + 1);
        const double integral_avg = fg / tv / fmx * c2[i];
        const double h = dp.lam * (fg + integral_avg);
        hprob_out[(size_t)blockIdx.x * periods + i] = h;
        if (f_out) f_out[i] = f[i];
        if (g_out) g_out[i] = g[i];
        if (i >= dp.start_indx && !(-1e-9 <= h && h <= 1.000000001)) bad = 1;
    }
    if (bad) atomicOr(&meta[blockIdx.x].status, PKB_ST_HPROB_RANGE);
}

// ---------------------------------------------------------------------------
struct PeriodInfo {
    double mux, muy;    // sub-cell remainder of the drift (cdf_mu, :485)
    int row_c, col_c;   // window centre in the domain grid (:491-494)
    int h;              // support half-width for this period
    int pad_;
};

// grid = problems, block = 256
__global__ void k_drift(const DayParams* __restrict__ dps, const BvnPar* __restrict__ bvn, const double* __restrict__ wind,
                        int periods, PeriodInfo* __restrict__ pinfo, DayMeta* __restrict__ meta, double ring_tol) {
    PKB_SHARED(double, red, 256);
    const DayParams dp = dps[blockIdx.x];
    const BvnPar& bp = bvn[dp.bvn_S];
    if (dp.kind) {
        // Day-0 spread kernel: integer / remainder split of the mean drift with Python's floor division and modulo
        // (Bayes_Run.py:247-251), support of the drifted blob (get_mvn_cdf_values with mu = remainder, :252-253), and
        // mlen//2 = max(h_long, h_short) + max(|xdrift_int|, |ydrift_int|) (:258-259) as the window extent.
        if (threadIdx.x == 0) {
            const double cell = dp.cell;
            const double fx = floor(dp.sdx / cell), fy = floor(dp.sdy / cell);
            PeriodInfo pi;
            pi.mux = dp.sdx - fx * cell;                  // x % res  (sign of the divisor, like Python)
            pi.muy = dp.sdy - fy * cell;
            pi.col_c = (int)fx;                           // xdrift_int
            pi.row_c = (int)fy;                           // ydrift_int
            const int h0 = bp.h0;
            int h = h0;
            while (h > 0 && 1.0 - square_prob(bp, cell, h - 1, pi.mux, pi.muy) < PKB_CDF_EPS) --h;
            while (!(1.0 - square_prob(bp, cell, h, pi.mux, pi.muy) < PKB_CDF_EPS) && h < 4096) ++h;
            const double d0 = 1.0 - square_prob(bp, cell, h, pi.mux, pi.muy);
            const double d1 = h >= 1 ? 1.0 - square_prob(bp, cell, h - 1, pi.mux, pi.muy) : 1.0;
            int flags = 0;
            if (fabs(d0 - PKB_CDF_EPS) < ring_tol || fabs(d1 - PKB_CDF_EPS) < ring_tol) {
                flags = PKB_ST_BORDERLINE;
                h = ring_halfwidth_ref_order(bp, cell, pi.mux, pi.muy, PKB_CDF_EPS, h + 1);
            }
            pi.h = h;
            pi.pad_ = 0;
            pinfo[(size_t)blockIdx.x * periods] = pi;
            const int hl = bvn[dp.bvn_Sl].h0;
            int ai = pi.col_c < 0 ? -pi.col_c : pi.col_c, aj = pi.row_c < 0 ? -pi.row_c : pi.row_c;
            meta[blockIdx.x].ext = (h > hl ? h : hl) + (ai > aj ? ai : aj);
            meta[blockIdx.x].hl = hl;
            if (flags) atomicOr(&meta[blockIdx.x].status, flags);
        }
        return;
    }
    const int P = dp.single ? 1 : periods;
    const double* w = wind + (size_t)dp.wind_day * periods * 3;
    const double* wn = w + (size_t)periods * 3;
    const int n = dp.n_periods;
    int ext = 0, flags = 0;
    for (int t = dp.start_indx + threadIdx.x; t < P; t += blockDim.x) {
        double mx, my;
        if (dp.single) {
            mx = w[0]; my = w[1];
        } else if (n > 1) {
            if (t + n - 1 < P) {
                double sx = 0.0, sy = 0.0;
                for (int i = 0; i < n; ++i) { sx += w[3 * (t + i)]; sy += w[3 * (t + i) + 1]; }
                mx = sx / n; my = sy / n;
            } else if (dp.has_next) {
                double sx, sy;
                if (t != P - 1) {
                    sx = 0.0; sy = 0.0;
                    for (int i = t; i < P; ++i) { sx += w[3 * i]; sy += w[3 * i + 1]; }
                } else { sx = w[3 * (P - 1)]; sy = w[3 * (P - 1) + 1]; }
                const int wrap = n - (P - t);
                if (wrap != 1) {
                    double ax = 0.0, ay = 0.0;
                    for (int i = 0; i < wrap; ++i) { ax += wn[3 * i]; ay += wn[3 * i + 1]; }
                    sx += ax; sy += ay;
                } else { sx += wn[0]; sy += wn[1]; }
                mx = sx / n; my = sy / n;
            } else {
                if (t != P - 1) {
                    double sx = 0.0, sy = 0.0;
                    for (int i = t; i < P; ++i) { sx += w[3 * i]; sy += w[3 * i + 1]; }
                    mx = sx / (P - t); my = sy / (P - t);
                } else { mx = w[3 * (P - 1)]; my = w[3 * (P - 1) + 1]; }
            }
        } else {
            mx = w[3 * t]; my = w[3 * t + 1];
        }
        const double scale = 86400.0 * ((double)n / (double)P);   // 3600*24*(n_periods/periods) (:468)
        mx *= scale; my *= scale;
        mx *= dp.mu_r; my *= dp.mu_r;                              // (:472)
        const double cell = dp.cell;
        PeriodInfo pi;
        pi.mux = mx - rint(mx / cell) * cell;                      // (:485)
        pi.muy = my - rint(my / cell) * cell;
        pi.col_c = dp.rad_res + (int)rint(mx / cell);              // (:491,494)
        pi.row_c = dp.rad_res + (int)rint(-my / cell);             // (:492,493)
        // support half-width: |cdf_mu| <= cell/2 => h in {h0-1, h0, h0+1}
        const int h0 = bp.h0;
        int h = h0 + 1;
        double d1 = 1.0, d0 = 1.0;
        if (h0 >= 1) d1 = 1.0 - square_prob(bp, cell, h0 - 1, pi.mux, pi.muy);
        d0 = 1.0 - square_prob(bp, cell, h0, pi.mux, pi.muy);
        if (h0 >= 1 && d1 < PKB_CDF_EPS) h = h0 - 1;
        else if (d0 < PKB_CDF_EPS) h = h0;
        if (fabs(d1 - PKB_CDF_EPS) < ring_tol || fabs(d0 - PKB_CDF_EPS) < ring_tol) {
            // too close to call on the one-rectangle form: the reference's own running sum decides
            flags |= PKB_ST_BORDERLINE;
            h = ring_halfwidth_ref_order(bp, cell, pi.mux, pi.muy, PKB_CDF_EPS, h0 + 1);
        }
        pi.h = h;
        pi.pad_ = 0;
        pinfo[(size_t)blockIdx.x * periods + t] = pi;
        int ro = pi.row_c - dp.rad_res, co = pi.col_c - dp.rad_res;
        if (ro < 0) ro = -ro;
        if (co < 0) co = -co;
        const int e = (ro > co ? ro : co) + h;
        if (e > ext) ext = e;
    }
    const int emax = (int)block_max((double)ext, red);
    if (flags) atomicOr(&meta[blockIdx.x].status, flags);
    if (threadIdx.x == 0) {
        meta[blockIdx.x].ext = emax;
        meta[blockIdx.x].hl = bvn[dp.bvn_Sl].h0;
    }
}

// python basic-slice length of seq[start:stop] for len n, start >= 0
__device__ __forceinline__ int py_slice_len(int start, int stop, int n) {
    if (stop < 0) { stop += n; if (stop < 0) stop = 0; }
    if (stop > n) stop = n;
    if (start > n) start = n;
    return stop > start ? stop - start : 0;
}

// ---------------------------------------------------------------------------
// grid = (periods, problems), block = 64 / 128 / 256 by lattice size (pkb200.cu), dyn smem = (6*nmax + tile_cap) doubles, tile_cap = min(PKB_LATTICE_CAP, nmax^2)
// acc: per problem (2*racc+1)^2 window centred on the release cell
#ifndef PKB_PERIOD_B
#define PKB_PERIOD_B 3
#endif
__global__ void __launch_bounds__(256, PKB_PERIOD_B) k_period(const DayParams* __restrict__ dps, const BvnPar* __restrict__ bvn, const PeriodInfo* __restrict__ pinfo,
                         const double* __restrict__ hprob, int periods, int nmax, int tile_cap, double* __restrict__ acc,
                         double* __restrict__ acc_lo, int racc, double* __restrict__ loss_t, DayMeta* __restrict__ meta) {
    PKB_DYN_SMEM(raw);
    PKB_SHARED(double, red, 256);
    PKB_SHARED(double, cstep, 20);
    const int prob = blockIdx.y;
    const int t = blockIdx.x;
    const DayParams dp = dps[prob];
    const int P = dp.single ? 1 : periods;
    if (dp.kind || t < dp.start_indx || t >= P) return;
    const BvnPar& bp = bvn[dp.bvn_S];
    const PeriodInfo pi = pinfo[(size_t)prob * periods + t];
    const double hp = hprob[(size_t)prob * periods + t];
    const int h = pi.h;
    const int n = 2 * h + 2;             // corners per side
    const int nc = 2 * h + 1;            // cells per side
    const int dom = 2 * dp.rad_res + 1;
    const int tid = threadIdx.x, T = blockDim.x;

    // --- window clipping exactly as the reference's index arithmetic (:501-527)
    int row_min = pi.row_c - h, row_max = pi.row_c + h;
    int col_min = pi.col_c - h, col_max = pi.col_c + h;
    int rs = 0, re = nc, cs = 0, ce = nc;
    if (row_max + 1 > dom) { re -= row_max + 1 - dom; if (re < 0) re = 0; row_max = dom - 1; }
    if (col_max + 1 > dom) { ce -= col_max + 1 - dom; if (ce < 0) ce = 0; col_max = dom - 1; }
    if (row_min < 0) { rs -= row_min; if (rs < 0) rs = 0; row_min = 0; }
    if (col_min < 0) { cs -= col_min; if (cs < 0) cs = 0; col_min = 0; }
    const int Ar = py_slice_len(row_min, row_max + 1, dom), Ac = py_slice_len(col_min, col_max + 1, dom);
    const int Br = py_slice_len(rs, re, nc), Bc = py_slice_len(cs, ce, nc);
    const bool shape_err = (Br != Ar && Br != 1) || (Bc != Ac && Bc != 1);   // numpy ValueError (:547)
    const bool clipped = rs > 0 || re < nc || cs > 0 || ce < nc;
    if (shape_err || Ar == 0 || Ac == 0) {
        // nothing lands in the domain: the whole period's mass is lost (:546,:558)
        if (tid == 0) {
            loss_t[(size_t)prob * periods + t] = hp;
            if (shape_err) atomicOr(&meta[prob].status, PKB_ST_WARNED);
        }
        return;
    }

    double* A = reinterpret_cast<double*>(raw);   // standardised x corners
    double* B = A + nmax;                         // standardised y corners
    double* PA = B + nmax;                        // Phi(-a)
    double* PB = PA + nmax;                       // Phi(-b)
    double* HA = PB + nmax;                       // a^2/2
    double* HB = HA + nmax;                       // b^2/2
    double* U = HB + nmax;                        // corner lattice tile
    const double cell = dp.cell, r = cell / 2;
    if (tid < 20) {
        const double dbs = cell / bp.sy;
        cstep[tid] = exp(-dbs * dbs * bp.inv[tid]);
    }
    for (int i = tid; i < n; i += T) {
        // corner i <= 2h is `low` of cell i - h; the last corner is `upp` of the last cell (:354-355)
        const double x = (i <= 2 * h) ? ((i - h) * cell - r) : ((h * cell - r) + cell);
        const double a = (x - pi.mux) / bp.sx, b = (x - pi.muy) / bp.sy;
        A[i] = a; B[i] = b;
        PA[i] = phid(-a); PB[i] = phid(-b);
        HA[i] = a * a / 2.0; HB[i] = b * b / 2.0;
    }
    __syncthreads();

    const int rows_per_tile = tile_cap / n - 1;          // cell rows (y) per tile
    const int shift = dp.rad_res - racc;                 // acc window origin in domain coords
    const int W = 2 * racc + 1;
    double* accp = acc + (size_t)prob * W * W;
    double* accl = acc_lo + (size_t)prob * W * W;
    double inside = 0.0;
    for (int y0 = 0; y0 < nc; y0 += rows_per_tile) {
        const int ny = (nc - y0 < rows_per_tile) ? (nc - y0) : rows_per_tile;   // cell rows in this tile
        const int npts = (ny + 1) * n;
        if (bp.high) {
            for (int q = tid; q < npts; q += T) {
                const int iy = q / n, ix = q - iy * n;
                U[q] = bvu_high(bp, A[ix], B[y0 + iy]);
            }
        } else {
            // |rho| < 0.925: every Gauss-Legendre node contributes w_i exp(E_i(a, b)) with
            //   E_i(a, b) = (sn_i a b - (a^2 + b^2) / 2) / (1 - sn_i^2),
            // a quadratic in b.  The corners of a lattice column are equally spaced in b, so along a
            // column exp(E_i) obeys T_{k+1} = T_k r_k, r_{k+1} = r_k c_i with the constant second
            // difference c_i = exp(-db^2 / (1 - sn_i^2)): a thread marches PKB_BVN_SEG corners of one
            // column with two exp() per node instead of one per node and corner (relative error
            // <= SEG^2 ulp on each term, i.e. ~1e-14 absolute on a cell mass against the 1e-10 bar).
            const int nrow = ny + 1;
            const int nseg = (nrow + PKB_BVN_SEG - 1) / PKB_BVN_SEG;
            const double db = cell / bp.sy;
            for (int it = tid; it < n * nseg; it += T) {
                const int seg = it / n, ix = it - seg * n;
                const int iy0 = seg * PKB_BVN_SEG;
                const int cnt = (nrow - iy0 < PKB_BVN_SEG) ? (nrow - iy0) : PKB_BVN_SEG;
                const double a = A[ix], b0 = B[y0 + iy0];
                const double hs0 = HA[ix] + HB[y0 + iy0];
                const double ab0 = a * b0, adb = a * db, lin = b0 * db + db * db / 2.0;
                double acc[PKB_BVN_SEG];
#pragma unroll
                for (int k = 0; k < PKB_BVN_SEG; ++k) acc[k] = 0.0;
                const int nn = 2 * bp.lg;
                for (int i = 0; i < nn; ++i) {
                    const double sn = bp.sn[i], inv = bp.inv[i], w = bp.w[i];
                    double Tk = exp(fma(sn, ab0, -hs0) * inv);
                    double rk = exp(fma(sn, adb, -lin) * inv);
                    const double c = cstep[i];
#pragma unroll
                    for (int k = 0; k < PKB_BVN_SEG; ++k) {
                        acc[k] = fma(w, Tk, acc[k]);
                        Tk *= rk;
                        rk *= c;
                    }
                }
                const double pa = PA[ix];
#pragma unroll
                for (int k = 0; k < PKB_BVN_SEG; ++k)
                    if (k < cnt) U[(iy0 + k) * n + ix] = fma(acc[k], bp.asr4pi, pa * PB[y0 + iy0 + k]);
            }
        }
        __syncthreads();
        const int ncell = ny * nc;
        // (row / column of a thread's cells advance incrementally: an integer division per cell was a fifth of this loop)
        int iy = tid / nc, ix = tid - iy * nc;
        const int dyT = T / nc, dxT = T - dyT * nc;
        for (int q = tid; q < ncell; q += T, iy += dyT, ix += dxT) {
            if (ix >= nc) { ix -= nc; ++iy; }
            const double* u = U + iy * n + ix;
            const double v = u[0] - u[1] - u[n] + u[n + 1];     // BVNMVN 4-term difference
            const int jj = y0 + iy - h;                          // y index (up)
            const int row = pi.row_c - jj;                       // rows run downwards (:377-378)
            const int col = pi.col_c + (ix - h);
            if (row >= 0 && row < dom && col >= 0 && col < dom) {
                inside += v;
                const size_t ci = (size_t)(row - shift) * W + (col - shift);
                acc_add_exact(&accp[ci], &accl[ci], hp * v);   // (:539)
            }
        }
        __syncthreads();
    }
    if (clipped) {
        const double s = block_sum(inside, red);
        if (tid == 0) loss_t[(size_t)prob * periods + t] = (1.0 - s) * hp;   // (:546)
    } else if (tid == 0) {
        loss_t[(size_t)prob * periods + t] = 0.0;
    }
}

// ---------------------------------------------------------------------------
// grid = problems x csize, block = 1024, launched as thread-block CLUSTERS of csize CTAs (csize = 1 or PKB_FIN_SLICES).
// Works in place on the accumulation window.  pre (optional): receives the pre-threshold window (parity export).
//
// The sub-window of a problem is cut into PKB_FIN_SLICES contiguous slices.  Every sum / minimum / maximum over the window is
// taken per slice (fixed tree inside the CTA) and the slice values are combined in slice order, so the result does not depend on
// how many CTAs share the work: with csize = 1 one CTA walks all slices (hundreds of problems in a launch: the likelihood batch),
// with csize = PKB_FIN_SLICES each CTA of the cluster takes one slice and the slice values travel through distributed shared
// memory (a handful of problems in a launch -- one solve's days -- would otherwise leave most SMs idle for several passes over
// 5 MB windows).
#define PKB_FIN_SLICES 8
struct FinCluster {
    int csize, crank;
#if !PKB_IS_EMUL
    __device__ __forceinline__ void sync() const { if (csize > 1) cooperative_groups::this_cluster().sync(); }
    // slot `s` of the array `part` in the CTA that owns slice s
    __device__ __forceinline__ const double* remote(double* part, int s) const {
        if (csize == 1) return part;
        return cooperative_groups::this_cluster().map_shared_rank(part, (unsigned)(s / (PKB_FIN_SLICES / csize)));
    }
#else
    void sync() const {}
    const double* remote(double* part, int) const { return part; }
#endif
};
// part[PKB_FIN_SLICES][3]: this CTA has filled the rows of its own slices; returns the slice-ordered combination of column k
// (op 0 sum, 1 min, 2 max) to every thread of every CTA of the cluster.  Two cluster barriers: after the fill, after the read.
__device__ __forceinline__ void fin_combine(const FinCluster& fc, double* part, int ncol, const int* op, double* out) {
    __syncthreads();
    fc.sync();
    for (int k = 0; k < ncol; ++k) {
        double r = op[k] == 0 ? 0.0 : (op[k] == 1 ? INFINITY : -INFINITY);
        for (int s = 0; s < PKB_FIN_SLICES; ++s) {
            const double v = fc.remote(part, s)[s * 3 + k];
            r = op[k] == 0 ? r + v : (op[k] == 1 ? fmin(r, v) : fmax(r, v));
        }
        out[k] = r;
    }
    fc.sync();
    __syncthreads();
}
__global__ void __launch_bounds__(1024) k_day_finalize(const DayParams* __restrict__ dps, const BvnPar* __restrict__ bvn, int periods, double* __restrict__ acc,
                               const double* __restrict__ acc_lo, int racc, const double* __restrict__ loss_t, DayMeta* __restrict__ meta,
                               double negval, double* __restrict__ pre, const PeriodInfo* __restrict__ pinfo, int csize) {
    PKB_SHARED(double, red, 1024);
    PKB_SHARED(double, sh, 4);
    PKB_SHARED(double, part, PKB_FIN_SLICES * 3);
    FinCluster fc;
    fc.csize = csize;
    fc.crank = (int)(blockIdx.x % (unsigned)csize);
    const int prob = (int)(blockIdx.x / (unsigned)csize);
    const DayParams dp = dps[prob];
    const int P = dp.single ? 1 : periods;
    const int W = 2 * racc + 1;
    const int nel = W * W;
    double* a = acc + (size_t)prob * nel;
    const int tid = threadIdx.x, T = blockDim.x;
    int status = 0;
    // The window is sized for the widest problem of the batch (racc); this problem's periods and its local blob only
    // reach max(ext, hl) cells from the release cell (k_drift), the rest of its window is still the zeros of the
    // initial memset.  All passes below walk that sub-window: element q -> (row, col) of the W x W window.
    const int hl = bvn[dp.bvn_Sl].h0;
    int rsub = meta[prob].ext > hl ? meta[prob].ext : hl;
    if (rsub > racc) rsub = racc;
    const int ns = 2 * rsub + 1, off = racc - rsub, nsub = ns * ns;
    auto at = [&](int q) -> int {
        const int rr = q / ns;
        return (rr + off) * W + (q - rr * ns + off);
    };
    if (dp.kind) {
        // Day-0 spread kernel (Bayes_Run.py:252-270): sprd = f * longsprd shifted by the integer drift, += (1 - f) *
        // shrtsprd at the centre, centre += max(0, 1 - sum).  No threshold, no renormalisation; its shape is mlen.
        // (a small blob: the first CTA of the cluster does it alone, no cluster barrier on this path)
        if (fc.crank != 0) return;
        const PeriodInfo pi = pinfo[(size_t)prob * periods];
        const BvnPar& bs = bvn[dp.bvn_S];
        const BvnPar& bl = bvn[dp.bvn_Sl];
        const double cell = dp.cell, r = cell / 2, f = dp.sprd_factor;
        const int h = pi.h, nc = 2 * h + 1;
        for (int q = tid; q < nc * nc; q += T) {
            const int iy = q / nc, ix = q - iy * nc;      // cell (x = ix - h, y = iy - h) of the drifted blob
            const double xl = (ix - h) * cell - r, yl = (iy - h) * cell - r;
            const double v = mvn_rect(bs, xl, xl + cell, yl, yl + cell, pi.mux, pi.muy);
            const int row = racc - ((iy - h) + pi.row_c), col = racc + ((ix - h) + pi.col_c);
            if (row >= 0 && row < W && col >= 0 && col < W) a[(size_t)row * W + col] = v * f;
        }
        __syncthreads();
        const int ncl = 2 * hl + 1;
        for (int q = tid; q < ncl * ncl; q += T) {
            const int iy = q / ncl, ix = q - iy * ncl;
            const double xl = (ix - hl) * cell - r, yl = (iy - hl) * cell - r;
            const double v = mvn_rect(bl, xl, xl + cell, yl, yl + cell, 0.0, 0.0);
            a[(size_t)(racc - (iy - hl)) * W + racc + (ix - hl)] += v * (1.0 - f);
        }
        __syncthreads();
        double s = 0.0, kc = 0.0;
        for (int q = tid; q < nsub; q += T) { const double v = a[at(q)]; s += v; kc += v != 0.0 ? 1.0 : 0.0; }
        const double tot = block_sum(s, red);
        const double cnt = block_sum(kc, red);
        if (pre) {
            for (int i = tid; i < nel; i += T) pre[(size_t)prob * nel + i] = a[i];
        }
        if (tid == 0) {
            const double fix = 1.0 - tot > 0.0 ? 1.0 - tot : 0.0;
            a[(size_t)racc * W + racc] += fix;
            DayMeta& m = meta[prob];
            m.loss = 0.0; m.pmfsum = tot; m.total = tot + fix; m.kept_sum = tot + fix; m.add = fix;
            m.rad = rsub; m.nnz = (int)cnt;
        }
        return;
    }

    // slices of this CTA: [s0, s1); slice s covers the elements [s * chunk, (s + 1) * chunk) of the sub-window
    const int per = PKB_FIN_SLICES / csize, s0 = fc.crank * per, s1 = s0 + per;
    const int chunk = (nsub + PKB_FIN_SLICES - 1) / PKB_FIN_SLICES;
    if (tid == 0) {
        double loss = 0.0;
        for (int t = dp.start_indx; t < P; ++t) loss += loss_t[(size_t)prob * periods + t];   // same order as the loop (:546,558)
        sh[0] = loss;
    }
    // exact two-plane accumulator (acc_add_exact) -> doubles, in place; sum and minimum in the same sweep
    const double* al = acc_lo + (size_t)prob * nel;
    for (int sl = s0; sl < s1; ++sl) {
        const int lo = sl * chunk, hi = lo + chunk < nsub ? lo + chunk : nsub;
        double s = 0.0, mn = 0.0;
        for (int q = lo + tid; q < hi; q += T) {
            const int i = at(q);
            const double v = acc_exact_to_double(a[i], al[i]);
            a[i] = v;
            s += v; mn = fmin(mn, v);
        }
        const double bs = block_sum(s, red), bm = block_min(mn, red);
        if (tid == 0) { part[sl * 3 + 0] = bs; part[sl * 3 + 1] = bm; }
    }
    double cmb[3];
    { const int op[2] = {0, 1}; fin_combine(fc, part, 2, op, cmb); }
    const double pmfsum = cmb[0], pmin = cmb[1];
    const double loss = sh[0];
    double total = pmfsum + loss;
    if (!(loss >= 0.0)) status |= PKB_ST_NEG_LOSS;
    if (!(pmin >= -1e-8)) status |= PKB_ST_PMF_NEG;
    if (!(pmfsum <= 1.00001)) status |= PKB_ST_PMF_GT1;
    if (total < 0.99999) {
        // wasps that did not fly diffuse locally around the release cell (:581-585)
        // The two assertions after the blob (:586-590) only need the new total and minimum.  The blob adds w * v to a few
        // cells: the total is the old one plus what was added, and with the old minimum above -1e-8 no cell can have dropped
        // below it unless a blob cell itself came out that negative -- so a second sweep over the window only runs when
        // the first assertion already failed (both statuses are errors for the caller either way).
        // (the blob is small: the first CTA of the cluster adds it; the cluster barrier in fin_combine publishes its cells)
        double added = 0.0, bmn = 0.0;
        if (fc.crank == 0) {
            const BvnPar& bl = bvn[dp.bvn_Sl];
            const int ncl = 2 * hl + 1;
            const double cell = dp.cell, r = cell / 2;
            const double wgt = 1.0 - total;
            for (int q = tid; q < ncl * ncl; q += T) {
                const int iy = q / ncl, ix = q - iy * ncl;      // iy: y index + hl, ix: x index + hl
                const double xl = (ix - hl) * cell - r, yl = (iy - hl) * cell - r;
                const double v = mvn_rect(bl, xl, xl + cell, yl, yl + cell, 0.0, 0.0);
                const int row = racc - (iy - hl), col = racc + (ix - hl);
                const double nv = a[(size_t)row * W + col] + wgt * v;
                a[(size_t)row * W + col] = nv;
                added += wgt * v; bmn = fmin(bmn, nv);
            }
            __threadfence();
        }
        {
            // (slot 0 belongs to the first CTA; the other slots carry neutral values)
            const double ba = block_sum(added, red), bb = block_min(bmn, red);
            for (int sl = s0; sl < s1; ++sl)
                if (tid == 0) { part[sl * 3 + 0] = sl == 0 ? ba : 0.0; part[sl * 3 + 1] = sl == 0 ? bb : 0.0; }
            const int op[2] = {0, 1};
            fin_combine(fc, part, 2, op, cmb);
        }
        double sum2 = pmfsum + cmb[0];
        double min2 = fmin(pmin, cmb[1]);
        if (!(pmin >= -1e-8)) {
            for (int sl = s0; sl < s1; ++sl) {
                const int lo = sl * chunk, hi = lo + chunk < nsub ? lo + chunk : nsub;
                double s = 0.0, mn = 0.0;
                for (int q = lo + tid; q < hi; q += T) { const double v = a[at(q)]; s += v; mn = fmin(mn, v); }
                const double bs = block_sum(s, red), bm = block_min(mn, red);
                if (tid == 0) { part[sl * 3 + 0] = bs; part[sl * 3 + 1] = bm; }
            }
            const int op[2] = {0, 1};
            fin_combine(fc, part, 2, op, cmb);
            sum2 = cmb[0];
            min2 = cmb[1];
        }
        if (!(min2 >= -1e-8)) status |= PKB_ST_PMF_NEG2;
        if (!(sum2 + loss <= 1.00001)) status |= PKB_ST_TOT_GT1;
    }
    // r_small_vals(coo(pmf), prob_model=True) (:605, CalcSol.py:112-136)
    for (int sl = s0; sl < s1; ++sl) {
        const int lo = sl * chunk, hi = lo + chunk < nsub ? lo + chunk : nsub;
        double ks = 0.0, kc = 0.0, kr = 0.0;
        for (int q = lo + tid; q < hi; q += T) {
            const int i = at(q);
            const double v = a[i];
            if (v != 0.0 && !(v < negval)) {
                ks += v; kc += 1.0;
                int rr = i / W - racc, cc = i % W - racc;
                if (rr < 0) rr = -rr;
                if (cc < 0) cc = -cc;
                kr = fmax(kr, (double)(rr > cc ? rr : cc));
            }
        }
        const double b0 = block_sum(ks, red), b1 = block_sum(kc, red), b2 = block_max(kr, red);
        if (tid == 0) { part[sl * 3 + 0] = b0; part[sl * 3 + 1] = b1; part[sl * 3 + 2] = b2; }
    }
    { const int op[3] = {0, 0, 2}; fin_combine(fc, part, 3, op, cmb); }
    const double ksum = cmb[0], kcnt = cmb[1], krad = cmb[2];
    const double add = (1.0 - ksum) / kcnt;
    if (pre) {      // parity export: the whole window, zeros outside the sub-window
        for (int i = fc.crank * T + tid; i < nel; i += T * csize) pre[(size_t)prob * nel + i] = 0.0;
        __syncthreads();
        fc.sync();
        for (int sl = s0; sl < s1; ++sl) {
            const int lo = sl * chunk, hi = lo + chunk < nsub ? lo + chunk : nsub;
            for (int q = lo + tid; q < hi; q += T) { const int i = at(q); pre[(size_t)prob * nel + i] = a[i]; }
        }
    }
    for (int sl = s0; sl < s1; ++sl) {
        const int lo = sl * chunk, hi = lo + chunk < nsub ? lo + chunk : nsub;
        for (int q = lo + tid; q < hi; q += T) {
            const int i = at(q);
            const double v = a[i];
            a[i] = (v != 0.0 && !(v < negval)) ? v + add : 0.0;
        }
    }
    if (tid == 0 && fc.crank == 0) {
        DayMeta& m = meta[prob];
        m.loss = loss; m.pmfsum = pmfsum; m.total = total; m.kept_sum = ksum; m.add = add;
        m.rad = (int)krad; m.nnz = (int)kcnt;
    }
    if (status && tid == 0 && fc.crank == 0) atomicOr(&meta[prob].status, status);
}

// ---------------------------------------------------------------------------
// get_mvn_cdf_values for arbitrary mu: grid = 1, block = 256.
// out: [cap] doubles, receives the (2h+1)^2 array in the reference orientation;
// hout[0] = h, or -1 if (2h+1)^2 > cap or h >= 4096.
__global__ void k_mvn_cdf(const BvnPar* __restrict__ bvn, double cell, double mux, double muy, double* __restrict__ out, int cap,
                          int* __restrict__ hout, double ring_tol) {
    PKB_SHARED(int, okv, 256);
    PKB_SHARED(int, found, 1);
    const BvnPar& p = bvn[0];
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) found[0] = -1;
    __syncthreads();
    for (int base = 0; base < 4096; base += T) {
        okv[tid] = (1.0 - square_prob(p, cell, base + tid, mux, muy) < PKB_CDF_EPS) ? 1 : 0;
        __syncthreads();
        if (tid == 0)
            for (int t = 0; t < T; ++t)
                if (okv[t]) { found[0] = base + t; break; }
        __syncthreads();
        if (found[0] >= 0) break;
    }
    if (found[0] >= 0) {
        // within ring_tol of cdf_eps the reference's running sum decides (see k_drift)
        const int hf = found[0];
        __syncthreads();
        if (tid == 0) {
            const double d0 = 1.0 - square_prob(p, cell, hf, mux, muy);
            const double d1 = hf >= 1 ? 1.0 - square_prob(p, cell, hf - 1, mux, muy) : 1.0;
            if (fabs(d0 - PKB_CDF_EPS) < ring_tol || fabs(d1 - PKB_CDF_EPS) < ring_tol)
                found[0] = ring_halfwidth_ref_order(p, cell, mux, muy, PKB_CDF_EPS, hf + 1);
        }
        __syncthreads();
    }
    const int h = found[0];
    const int nc = 2 * h + 1;
    if (h < 0 || (long long)nc * nc > cap) { if (tid == 0) hout[0] = -1; return; }
    const double r = cell / 2;
    for (int q = tid; q < nc * nc; q += T) {
        const int row = q / nc, col = q - row * nc;
        const int jj = h - row, ii = col - h;                   // [row, col] = (y = h - row, x = col - h) (:377-378)
        const double xl = ii * cell - r, yl = jj * cell - r;
        out[q] = mvn_rect(p, xl, xl + cell, yl, yl + cell, mux, muy);
    }
    if (tid == 0) hout[0] = h;
}

}  // namespace pkb
