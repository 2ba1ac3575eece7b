/* pkb200.h -- C ABI of libpkb200.so, the B200 (sm_100a) implementation of the
 * Parasitoids drift-diffusion forward solve.
 *
 * The reference (mountaindust/Parasitoids) is pure Python and has no FFI of its
 * own; its plugin seam for this path is the Python class cuda_lib.CudaSolve
 * (cuda_lib.py:16-221) plus the module-level functions of ParasitoidModel.py
 * and CalcSol.py.  Each entry point below names the reference interface it
 * replaces; parasitoids_b200/*.py binds them with ctypes and re-exposes the
 * reference's Python names.
 *
 * Conventions: every function returns 0 on success or a negative PKB_E* code
 * and stores a message retrievable with pkb_last_error() (thread-local).  All
 * pointers are caller-owned HOST buffers of C-order IEEE fp64 / int32 unless a
 * comment says otherwise.  Handles are not thread-safe; use one per thread.
 * There is no CPU fallback: without a CUDA device pkb_create() fails.
 */
#ifndef PKB200_H
#define PKB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PKB_OK 0
#define PKB_EINVAL (-1)   /* bad argument */
#define PKB_ECUDA (-2)    /* CUDA runtime error */
#define PKB_ENOMEM (-3)
#define PKB_ELIMIT (-4)   /* problem exceeds an implementation limit */
#define PKB_ESTATE (-5)   /* call sequence error */

/* status bits of pkb_day_meta.status (reference assertion / warning sites) */
#define PKB_ST_HPROB_RANGE 1   /* ParasitoidModel.py:529 */
#define PKB_ST_NEG_LOSS 2      /* :569 */
#define PKB_ST_PMF_NEG 4       /* :570 */
#define PKB_ST_PMF_GT1 8       /* :571 */
#define PKB_ST_PMF_NEG2 16     /* :589 */
#define PKB_ST_TOT_GT1 32      /* :590 */
#define PKB_ST_WARNED 64       /* RuntimeWarning, :547-558 */
#define PKB_ST_BORDERLINE 128  /* ring-growth test (:348) within ring_tol of cdf_eps: decided by the reference-order running sum */

typedef struct pkb_ctx pkb_ctx;       /* one device + stream + plan cache */
typedef struct pkb_kset pkb_kset;     /* device-resident set of per-day kernels */
typedef struct pkb_chain pkb_chain;   /* convolution-chain state (CudaSolve instance) */
typedef struct pkb_result pkb_result; /* outputs of a fused solve */

/* arguments of one prob_mass evaluation (ParasitoidModel.py:384-385) */
typedef struct pkb_day_args {
    double hparams[7];   /* lam, aw, bw, a1, b1, a2, b2 */
    double dparams[3];   /* sig_x, sig_y, rho   (in-flow diffusion) */
    double dlparams[3];  /* sig_x, sig_y, rho   (out-of-flow diffusion) */
    double mu_r;
    double rad_dist;
    double start_time;   /* fraction of the day; < 0 means None */
    int n_periods;
    int rad_res;
    int wind_day;        /* row of the wind array that holds this day */
    int single;          /* 1: wind row is a single (wx, wy, wr) triple (test form, :426-428) */
    /* kind 1: not a prob_mass day but the local day-0 spread kernel the Bayes drivers prepend when the wind record
     * starts a day late (Bayes_Run.py:245-270, Bayes_MAP.py:247-277): sprd_factor * get_mvn_cdf_values(dparams)
     * shifted by the integer part of sprd_drift, plus (1 - sprd_factor) * get_mvn_cdf_values(dlparams), centre
     * topped up to unit mass; not thresholded.  Only dparams, dlparams, rad_dist and rad_res are read besides. */
    int kind;
    int pad_;
    double sprd_factor;
    double sprd_drift[2];   /* mean drift in metres (the reference uses (-25, 15)) */
} pkb_day_args;

typedef struct pkb_day_meta {
    double loss, pmfsum, total, kept_sum, add;
    int rad;      /* returned pmf has shape (2*rad+1)^2 */
    int nnz;
    int status;   /* PKB_ST_* bits */
    int ext, hl, pad_;
} pkb_day_meta;

typedef struct pkb_step_meta {
    double padmax, ksum, add, padabs;   /* padmax: max over the pad (the flag compares it with 1e-8); padabs: max |value| there */
    long long kcnt;
    int flag;     /* boundary flag, CalcSol.py:36-40 */
    int spec;     /* this step started from the stored spectrum of the state (option "spectral") */
    int wr0, wr1; /* the step only computed the rows [wr0, wr1) -- the others hold nothing above 1e-15 (support-window
                   * steps, option "tau_windows"; spectral-resident steps, option "spectral_rows"); wr1 <= wr0: all rows */
    int wc0, wc1; /* ... and the columns [wc0, wc1) (support-window steps) */
    int er0, er1, ec0, ec1;   /* measured extent [er0, er1) x [ec0, ec1) of the cells >= 1e-15 of the day's state */
} pkb_step_meta;

const char* pkb_last_error(void);
int pkb_version(void);

/* ---- context ------------------------------------------------------------- */
int pkb_create(int device, pkb_ctx** out);
int pkb_destroy(pkb_ctx* ctx);
int pkb_sync(pkb_ctx* ctx);
/* option keys: "stencil_max_radius" (direct-convolution switch point), "fft_threads",
 * "windows" (0/1: support-window chain steps in pkb_solve, default 1),
 * "fuse_rows" (0/1: inverse row pass also runs the next step's forward row pass, default 1),
 * "step_torus" (0/1: whole-torus steps on the smallest 7-smooth torus >= P + 2m of that day's kernel, default 1),
 * "trunc_torus" (0/1: steps from a truncated (flagged) state on a torus >= dom_len + 2m, default 1),
 * "spectral" (0/1: spectral-resident chain steps while the content outside the domain is below 1e-13, default 1;
 *             the one option whose results differ by more than rounding: by at most 1e-11, see chain.cuh),
 * "tau_windows" (0/1: support-window steps follow the numerical support -- cells >= 1e-15 -- instead of the exact one,
 *                default 1; "tau_lag": they use the extent measured this many steps back, default 1),
 * "spectral_rows" (0/1: such a step inverse-transforms only the rows that can hold a cell above 1e-15, default 1),
 * "spectral_min_reach" (arm them only if the exact support stays inside the domain for this many steps, default 4),
 * "ring_tol" (support-ring decisions of get_mvn_cdf_values closer than this to cdf_eps are re-taken with the reference's
 *             own running sum, ParasitoidModel.py:345-373; default 1e-12, 1.0 forces that path everywhere),
 * "batch_lanes" (1..8: proposals of pkb_solve_batch in flight at once, default 4),
 * "batch_group" (proposals per kernel-construction group of pkb_solve_batch, default 32),
 * "batch_chain" (0/1: the chains of a group run as batched kernels -- step n of every proposal in one launch per pass,
 *               csrc/bchain.cuh -- instead of one chain per proposal on the lanes, default 1; probability model and
 *               one-day releases, the other proposals take the per-proposal path either way),
 * "batch_occ" (resident CTAs per SM the batched chain kernels are launched for, default 4),
 * "fin_clusters" (0/1: k_day_finalize as a thread-block cluster of eight CTAs per (proposal, day) problem when a launch carries
 *                only a few dozen problems, default 1; same bits as one CTA per problem),
 * "cohort_lanes" (0/1: population model with a release of several days -- the cohort back-solves and the emission of day n run
 *                on child contexts next to the main chain's step n + 1, default 1),
 * "coo_thread" (0/1: the per-day COO / CSR compaction and D2H of pkb_solve are enqueued by a helper host thread instead of
 *              the thread that paces the chain, default 1).
 * All of them select between implementations of the same arithmetic; results agree to rounding. */
int pkb_set_option(pkb_ctx* ctx, const char* key, double value);
/* device time in ms of the phases of the last pkb_solve: [0] phase 1,
 * [1] chain, [2] output compaction + D2H, [3] total */
int pkb_timing(pkb_ctx* ctx, double out_ms[4]);
/* number of kernel launches issued through this context so far */
long long pkb_launch_count(pkb_ctx* ctx);
/* CUDA-event stopwatch on the context's stream: record mark `slot` (0..7);
 * device time between two recorded marks (waits for the later one) */
int pkb_mark(pkb_ctx* ctx, int slot);
int pkb_elapsed_ms(pkb_ctx* ctx, int slot_a, int slot_b, double* ms);
/* per-kernel device timing: while enabled every launch is bracketed by CUDA
 * events on the context's stream; pkb_profile_get returns the launch count and
 * summed duration of one kernel by name (e.g. "k_cols") since the last reset */
int pkb_profile_enable(pkb_ctx* ctx, int on);
int pkb_profile_reset(pkb_ctx* ctx);
int pkb_profile_get(pkb_ctx* ctx, const char* kernel, long long* count, double* total_ms);

/* ---- phase 1: ParasitoidModel.py ------------------------------------------ */
/* The interpolation of get_wind_data(site_name, interp_num, start_time)  (ParasitoidModel.py:162-227).
 * raw: [nd][npts][3] (wx, wy, wr) of consecutive days as read from the wind file; out: [nd][npts*interp_num][3].
 * half_hour_start: 0 for start_time '00:00', 1 for '00:30'. */
int pkb_wind_interp(pkb_ctx* ctx, const double* raw, int nd, int npts, int interp_num, int half_hour_start, double* out);

/* h_flight_prob(day_wind, lam, aw, bw, a1, b1, a2, b2)  (ParasitoidModel.py:282-309)
 * wind: [periods][3] (or [3] when single != 0); out: [periods] (or [1]) */
int pkb_hprob(pkb_ctx* ctx, const double* wind, int periods, int single, const double hparams[7], double* out,
              double* f_out /* f_time_prob, :243-267, may be NULL */, double* g_out /* g_wind_prob, :231-240, may be NULL */);

/* get_mvn_cdf_values(cell_length, mu, S)  (ParasitoidModel.py:311-380)
 * cov = (S[0][0], S[1][1], S[0][1]); out receives (2h+1)^2 values, h_out the
 * half-width; fails with PKB_ELIMIT if (2h+1)^2 > cap */
int pkb_mvn_cdf(pkb_ctx* ctx, double cell_length, const double mu[2], const double cov[3], double* out, int cap, int* h_out);

/* prob_mass for `nprob` independent (proposal, day) problems sharing one wind
 * array [nd_wind][periods][3]  (ParasitoidModel.py:384-613; the fan-out is
 * Run.py:412-425 / Bayes_Run.py:236-272).  All problems must share rad_res.
 * keep_pre != 0 also keeps the pre-threshold grids (parity export). */
int pkb_kernels_build(pkb_ctx* ctx, const double* wind, int nd_wind, int periods, const pkb_day_args* args, int nprob,
                      int keep_pre, pkb_kset** out);
int pkb_kset_meta(pkb_kset* ks, int i, pkb_day_meta* out);
/* dense (2*rad+1)^2 thresholded + renormalised pmf of problem i */
int pkb_kset_get(pkb_kset* ks, int i, double* out);
/* dense (2*racc+1)^2 pre-threshold window (needs keep_pre); racc via pkb_kset_racc */
int pkb_kset_get_pre(pkb_kset* ks, int i, double* out);
int pkb_kset_racc(pkb_kset* ks);
/* per-period (row_cent, col_cent, h) triples and hprob of problem i */
int pkb_kset_periods(pkb_kset* ks, int i, int* rch /*[periods][3]*/, double* hprob /*[periods]*/);
int pkb_kset_destroy(pkb_kset* ks);

/* ---- phase 2: CalcSol.py / cuda_lib.CudaSolve ----------------------------- */
/* CudaSolve.__init__(A, max_shape)  (cuda_lib.py:18-54; CalcSol.fft2, CalcSol.py:11-24)
 * dom_len: side of the square domain; max_shape: side of the largest filter.
 * The reference torus is P = dom_len + max_shape/2. */
int pkb_chain_create(pkb_ctx* ctx, int dom_len, int max_shape, pkb_chain** out);
int pkb_chain_destroy(pkb_chain* ch);
/* state := A (dense dom_len^2), zero padded */
int pkb_chain_set_state(pkb_chain* ch, const double* A);
/* state := kernel i of ks re-centred on the domain (Run.py:454-458) */
int pkb_chain_set_state_kernel(pkb_chain* ch, pkb_kset* ks, int i);
/* CudaSolve.fftconv2(B)  (cuda_lib.py:58-94; CalcSol.fftconv2, CalcSol.py:45-66)
 * B: dense odd-sided square (k x k), centre at [k/2][k/2] */
int pkb_chain_conv(pkb_chain* ch, const double* B, int k);
int pkb_chain_conv_kernel(pkb_chain* ch, pkb_kset* ks, int i);
/* CudaSolve.get_cursol(dom_shape, negval)  (cuda_lib.py:98-140; CalcSol.ifft2 +
 * the flag/re-fft logic of CalcSol.py:197-201).  apply_trunc != 0 applies the
 * boundary-flag truncation to the chain state (the "Re-fft" of cuda_lib.py:130-136);
 * apply_trunc == 0 only reports the flag (CalcSol.ifft2).  mode 0: raw domain
 * values (CPU-path ifft2), 1: entries < negval zeroed, 2: r_small_vals(
 * prob_model=True) applied.  out: dense dom_len^2 (may be NULL). */
int pkb_chain_get_cursol(pkb_chain* ch, double negval, int mode, int apply_trunc, double* out, pkb_step_meta* meta);
/* CudaSolve.back_solve(prev_spread, dom_shape)  (cuda_lib.py:145-221;
 * CalcSol.back_solve, CalcSol.py:72-109 with the same-shape re-FFT).
 * filters: nf dense k_i x k_i arrays in chronological order; out: nf dense
 * dom_len^2 grids in emergence order (may be NULL: cohorts stay on the device
 * for pkb_chain_population); threshold < 0: un-thresholded (CPU path), else
 * entries <= threshold zeroed per cohort (cuda_lib path). */
int pkb_chain_back_solve(pkb_chain* ch, const double* const* filters, const int* ks, int nf, double threshold, double* out,
                         int* flags);
/* Cohort superposition of the population model (CalcSol.py:236-237,271-274,
 * 303-306,322-323): out = r_small_vals((sum_c cohort_c * weights[c]) * r_number)
 * (+ centre_extra at the release cell if add_centre).  Cohorts 0..ncoh-2 are
 * those left on the device by the last pkb_chain_back_solve, cohort ncoh-1 is
 * the current state.  first_day != 0: the day-0 form r_small_vals(state) *
 * r_number * weights[0].  out (and optional pre, the un-thresholded sum):
 * dense dom_len^2. */
int pkb_chain_population(pkb_chain* ch, int ncoh, const double* weights, double r_number, double centre_extra, int add_centre,
                         double negval, int first_day, double* out, double* pre);
/* full padded P x P state (diagnostics / tests); P via pkb_chain_dims */
int pkb_chain_get_state(pkb_chain* ch, double* out);
int pkb_chain_dims(pkb_chain* ch, int* D, int* P, int* N);

/* ---- diagnostics ---------------------------------------------------------- */
/* length-n complex DFT of one vector through the shared-memory FFT used by the
 * chain (interleaved re,im; natural order in and out; inverse is unnormalised).
 * n must be 7-smooth.  Test hook: the reference reaches pocketfft through
 * scipy.fftpack (CalcSol.py:24,35). */
int pkb_debug_fft(pkb_ctx* ctx, int n, const double* in, double* out, int inverse);
/* smallest 7-smooth length >= n (the torus side the chain uses) */
int pkb_smooth_len(int n);

/* ---- fused forward solve: Run.main's hot path (Run.py:399-481) ------------- */
typedef struct pkb_solve_args {
    const double* wind;      /* [nd_wind][periods][3] */
    int wind_on_device;      /* wind is a device pointer (already resident) */
    int nd_wind, periods;
    int ndays;               /* days to simulate (<= nd_wind) */
    pkb_day_args day;        /* model parameters; wind_day/start_time are per-day and filled internally */
    int prob_model;          /* 1: get_solutions, 0: get_populations */
    int r_dur;               /* population model: release duration (days) */
    double r_number;
    const double* r_dist;    /* [r_dur] emergence fractions dist(1..r_dur) */
    double r_start;          /* start_time of day 0 for the population model; < 0 none */
    double negval;           /* 1e-8 */
    int want_dense_host;     /* copy dense solutions to host */
    int want_coo;            /* 1: build COO (row-major) on device and copy to host (pkb_result_coo); 2: the same triplets as CSR --
                              * per-row offsets instead of a row index per non-zero, 12 instead of 16 bytes each (pkb_result_csr;
                              * the layout Run.main saves, Run.py:490-510) */
    int keep_dense_device;   /* keep [ndays][D][D] on device (bench / gather) */
    int sprd;                /* 1: prepend the day-0 spread kernel (pkb_day_args kind 1) built from day.dparams / day.dlparams,
                              * run the chain over ndays + 1 days and drop the first (Bayes_Run.py:245-296) */
    double sprd_factor;
    double sprd_drift[2];
    const double* sprd_factors; /* pkb_solve_batch only: one sprd_factor per proposal (it is a sampled variable of its own,
                                 * Bayes_Run.py:202); NULL: sprd_factor for all */
    int out_on_device;       /* pkb_solve_batch(_projected) only: `out` is DEVICE memory (e.g. the tensor the caller all-gathers
                              * over NCCL): the results never visit the host */
    int keep_pre_device;     /* parity export: also keep every day's UN-thresholded domain grid (the `A[:D,:D]` of
                              * CalcSol.py:189-190 / the cohort sum of :322 before r_small_vals) for pkb_result_pre */
} pkb_solve_args;

int pkb_solve(pkb_ctx* ctx, const pkb_solve_args* args, pkb_result** out);
/* Likelihood batch (the call pattern of Bayes_Run.py:204-336, one forward solve per MCMC proposal, for
 * `nprop` proposals at once): proposals[nprop][15] in the order of the block-updated variables
 * (Bayes_Run.py:186-187: g_aw g_bw f_a1 f_b1 f_a2 f_b2 sig_x sig_y corr sig_x_l sig_y_l corr_l lam
 * n_periods mu_r); everything else (wind, domain, release settings) from `base`.  out[nprop][ndays][K]
 * receives the model at the K (row, col) sample cells -- what popdensity_to_emergence / popdensity_grid
 * read (Bayes_funcs.py:58-74,167-173); status[nprop][ndays] (may be NULL) the PKB_ST_* bits of every
 * (proposal, day) kernel.  Kernel construction is batched over groups of proposals, and so are the chains: step n of
 * every proposal of a group is one launch per pass (csrc/bchain.cuh). */
int pkb_solve_batch(pkb_ctx* ctx, const pkb_solve_args* base, const double* proposals, int nprop, const int* cells /*[K][2]*/, int K,
                    double* out, int* status);
/* ---- likelihood projection: Bayes_funcs.py:20-180 on the device --------------------------------------------
 * popdensity_to_emergence / popdensity_grid read the model only at a few cells and fold it with the incubation
 * distribution; both are an ordered linear map of the model at K sample cells, described by the caller (the
 * Python side builds it from a LocInfo exactly as the reference's loops run, parasitoids_b200/Bayes_funcs.py):
 *   S[day][set]   = numpy-sum of the model over the sample cells set_cells[set_ptr[set] .. set_ptr[set+1])
 *   G[group]      = ((0 + S[term_day][term_set] * term_w) + ...) over the terms grp_ptr[group] .. grp_ptr[group+1)
 *   out[row]      = numpy-sum of G over the groups row_ptr[row] .. row_ptr[row+1)
 * ("numpy-sum": numpy's pairwise summation order, restated in csrc/project.cuh.)  All arrays are host arrays. */
typedef struct pkb_projection {
    int nsets;
    const int* set_ptr;     /* [nsets + 1] */
    const int* set_cells;   /* indices into the K sample cells */
    int nrows;
    const int* row_ptr;     /* [nrows + 1] -> groups */
    int ngroups;
    const int* grp_ptr;     /* [ngroups + 1] -> terms */
    const int* term_day;    /* model day (0 = first output day) */
    const int* term_set;
    const double* term_w;
} pkb_projection;
/* out[nprop][nrows] from host samples[nprop][ndays][K] (what pkb_solve_batch returns) */
int pkb_project(pkb_ctx* ctx, const pkb_projection* proj, const double* samples, int nprop, int ndays, int K, double* out);
/* pkb_solve_batch with the projection applied on the device: out[nprop][proj->nrows]; only the projected values
 * cross PCIe (Bayes_Run.py:298-306: popdensity_to_emergence + popdensity_grid straight after get_populations) */
int pkb_solve_batch_projected(pkb_ctx* ctx, const pkb_solve_args* base, const double* proposals, int nprop, const int* cells /*[K][2]*/,
                              int K, const pkb_projection* proj, double* out, int* status);
/* ---- one solve over the G GPUs of a box (one process per GPU) ----------------------
 * Phase 1: every rank builds the kernels of its share of the days (the fan-out of Run.py:422-425), exports them into
 * a caller-owned device buffer, the caller all-gathers those (NCCL) and hands the complete set back as a kernel set.
 * Phase 2: the slab-decomposed spectral-resident chain of csrc/dist.cuh; the caller drives one day at a time and
 * runs the two collectives between the passes on the same stream:
 *     pkb_dist_step_cols(day)  ->  all_to_all(send -> recv)  ->  pkb_dist_step_rows()
 *                              ->  all_gather(stats -> allstats)  ->  pkb_dist_step_emit(day)
 * parasitoids_b200/multi.py does exactly that with torch.distributed. */
typedef struct pkb_dist pkb_dist;
/* launch everything of this context on `stream` (a cudaStream_t owned by the caller, e.g. torch's current stream);
 * NULL restores the context's own stream */
int pkb_set_stream(pkb_ctx* ctx, void* stream);
/* centred (Wdst x Wdst) window of kernel i, zero padded, into caller-owned DEVICE memory */
int pkb_kset_export_device(pkb_kset* ks, int i, void* dst_dev, int Wdst);
/* kernel set from n (W x W) windows in DEVICE memory (copied), with their crop radii */
int pkb_kset_from_device(pkb_ctx* ctx, const void* windows_dev, int n, int W, const int* rads, int rad_res, pkb_kset** out);
/* geometry for `world` ranks: complex elements of the send (= recv) buffer, rows per rank, torus */
int pkb_dist_plan(pkb_ctx* ctx, int dom_len, int mmax, int world, long long* xchg_elems, int* rows_per_rank, int* P, int* N);
/* send / recv: xchg_elems complex each; stats: 4 doubles; allstats: 4 * world doubles; out: [ndays][rows_per_rank][dom_len]
 * doubles (this rank's rows of every day's thresholded + renormalised solution; rows beyond the domain are zero).
 * All caller-owned DEVICE memory.  Day 0 (the recentred first kernel, Run.py:454-458) is written at once. */
int pkb_dist_create(pkb_ctx* ctx, pkb_kset* ks, int ndays, int rank, int world, void* send, void* recv, void* stats, void* allstats,
                    void* out, pkb_dist** handle);
int pkb_dist_step_cols(pkb_dist* h, int day);
int pkb_dist_step_rows(pkb_dist* h);
int pkb_dist_step_emit(pkb_dist* h, int day);
/* after the last day (synchronises): meta[ndays][4] = (kept sum, kept count, max outside the domain, max |.| outside)
 * per day; ok = 0 if the criterion for skipping the fold mod P failed on some day (then the results must be
 * discarded and the solve repeated with pkb_solve) */
int pkb_dist_finish(pkb_dist* h, double* meta, int* ok);
int pkb_dist_destroy(pkb_dist* h);

int pkb_result_info(pkb_result* r, int* ndays, int* dom_len, int* P, int* N, int* max_shape);
/* number of chain steps that ran on a support-window torus smaller than N (exact: the state is
 * identically zero outside the window while the spread has not reached the domain edge) */
int pkb_result_window_steps(pkb_result* r, int* n);
int pkb_result_day_meta(pkb_result* r, int day, pkb_day_meta* kmeta, pkb_step_meta* smeta);
/* population model with r_dur > 1: flag / sums of the back_solve step (CalcSol.py:99-105) that produced cohort
 * `cohort` (0 = first release day) on `day`; zeros when that cohort was not back-solved on that day */
int pkb_result_cohort_meta(pkb_result* r, int day, int cohort, pkb_step_meta* smeta);
/* dense solution of one day: device->host copy on demand */
int pkb_result_dense(pkb_result* r, int day, double* out);
/* dense UN-thresholded grid of one day (needs keep_pre_device) */
int pkb_result_pre(pkb_result* r, int day, double* out);
/* COO of all days (host pointers owned by the result, valid until destroy) */
int pkb_result_coo(pkb_result* r, const long long** day_offsets /*[ndays+1]*/, const int** rows, const int** cols,
                   const double** vals);
/* CSR of all days (want_coo = 2): triplets of day d are [day_offsets[d], day_offsets[d+1]) of cols / vals; row r of day d starts
 * at day_offsets[d] + row_offsets[d * dom_len + r] */
int pkb_result_csr(pkb_result* r, const long long** day_offsets /*[ndays+1]*/, const long long** row_offsets /*[ndays][dom_len]*/,
                   const int** cols, const double** vals);
/* gather values at K (row, col) cells for every day: out[ndays][K] */
int pkb_result_sample(pkb_result* r, const int* cells /*[K][2]*/, int K, double* out);
/* projection of one solve's dense device-resident days (keep_dense_device / want_dense_host): out[proj->nrows] */
int pkb_result_project(pkb_result* r, const int* cells /*[K][2]*/, int K, const pkb_projection* proj, double* out);
/* device pointer of the dense solutions [ndays][D][D] (keep_dense_device) */
int pkb_result_device_ptr(pkb_result* r, void** dptr);
int pkb_result_destroy(pkb_result* r);

#ifdef __cplusplus
}
#endif
#endif /* PKB200_H */
