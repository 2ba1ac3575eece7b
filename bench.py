#!/usr/bin/env python
"""Benchmark of the drift-diffusion forward solve (BASELINE.json metric:
simulated days/sec, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun ... bench.py --gpus N ...        (N > 1, one rank per GPU)

One *step* = one complete forward solve of the synthetic workload (SURVEY.md
section 8d, config C4): 60 flight days on a 4097 x 4097 domain, hourly wind
interpolated to 1440 take-off periods per day -- per-day kernel construction
plus the 59-step convolution chain.

`value`     days/s with the wind series already resident in HBM and the dense
            daily solutions left on the device (device-side work only).
`e2e`       the same solve through the public API (parasitoids_b200.Run.solve)
            with HOST buffers: wind uploaded from host memory and the
            thresholded CSR arrays of every day copied back inside the timed
            region -- what a caller of Run.main sees.
`roofline`  the dominant chain kernel against the measured HBM copy bandwidth;
`roofline_chain` all chain kernels of one simulated day against B_day(P, D).
`roofline_phase1` kernel construction (k_period) against the measured fp64 DFMA
            and exp rates; executed flops from the committed ncu capture.
`cpu_baseline` the CPU path (oracle/, the numpy restatement of the reference)
            on a bounded sample, phase 1 fanned out over a multiprocessing
            pool like Run.py:422-425.
`--impl reference` the UNMODIFIED reference (ParasitoidModel.prob_mass with the
            3-line mvnun shim, CalcSol.get_solutions) from baseline/_ref on a
            bounded sample (kind "reference"); the oracle port if that copy is
            absent (kind "port").
`extra`     the other BASELINE.json configurations at every N: c5 (512 prior
            draws x the Kalbar population solve through batch.solve_batch,
            strong scaling), weak_batch (one C4 proposal per rank through the
            same solve_batch path, weak scaling); at N = 1 also c1/c2/c3
            (Kalbar / Carnarvon at 801^2) with the CPU port beside them.

N > 1: the SAME step on every rank -- one independent solve per GPU through
Run.solve, exactly the N = 1 code path (replicas: one solve does not shard
efficiently over GPUs; `--single-solve` measures that).
The time is the maximum over the ranks between two barriers; no data-path
collective.  The batched-likelihood partitioning of SURVEY.md section 8e, with
its NCCL all_gather, is measured over the same ranks under `extra`
(weak_batch, c5).  `--workload kalbar_batch512` makes the likelihood batch the
step itself.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HPARAMS = (1., 1.263, 3.913, 7.302, 2.614, 23.999, 2.350)      # Run.py:68-83
DPARAMS = (171.82, 144.58, 0.253)
DLPARAMS = (7.096, 7.260, 0.000)
MU_R = 1.179
N_PERIODS = 30

WORKLOADS = {
    # name: (days, hourly samples/day, interp_num, rad_dist, rad_res)
    'synthetic_4097x4097_60d': (60, 24, 60, 51200.0, 2048),
    'synthetic_801x801_18d': (18, 24, 60, 10000.0, 400),          # Kalbar-sized, for quick checks
    # BASELINE.json configs[4]: 512 MCMC proposals x the full Kalbar population solve (Run.py:126-138),
    # sharded over the ranks by parasitoids_b200.batch.solve_batch (strong scaling)
    'kalbar_batch512': (18, 48, 30, 10000.0, 400),
}
BATCH = 512


SITES = {'kalbar': ('kalbar', '00:00'), 'carnarvon': ('carnarvonearl', '00:30')}       # Run.py:108-138


def _wind_reader(cpu):
    """get_wind_data: the package's (interpolation on the device) or, for the CPU arms, the oracle's restatement."""
    if cpu:
        from oracle import pm_oracle as PO
        return PO.get_wind_data
    from parasitoids_b200 import ParasitoidModel as PM
    return PM.get_wind_data


def stack_wind(wind_data, days):
    return np.stack([np.asarray(wind_data[d], dtype=float) for d in days])


def site_wind(site, cpu=False):
    """data/<site>wind.txt (the reference's wind records, shipped under data/) through get_wind_data."""
    name, start = SITES[site]
    wind_data, days = _wind_reader(cpu)(os.path.join(ROOT, 'data', name), 30, start)
    return stack_wind(wind_data, days), wind_data, days, 10000.0, 400


def prior_proposals(B, seed=7):
    """B draws from the central region of the priors of Bayes_Run.py:102-130 (PyMC2
    alpha/beta and tau parameterisations), in batch.PROPOSAL_FIELDS order."""
    rng = np.random.default_rng(seed)
    sd = 1.0 / np.sqrt(0.3)
    P = np.empty((B, 15))
    P[:, 0] = rng.gamma(2.2, 1.0, B)                                   # g_aw
    P[:, 1] = rng.gamma(5.0, 1.0, B)                                   # g_bw
    P[:, 2] = np.clip(rng.normal(6.0, sd, B), 0.0, 9.0)                # f_a1
    P[:, 3] = 1.0 + rng.gamma(2.0, 1.0, B)                             # f_b1
    P[:, 4] = np.clip(rng.normal(20.0, sd, B), 15.0, 24.0)             # f_a2
    P[:, 5] = 1.0 + rng.gamma(2.0, 1.0, B)                             # f_b2
    P[:, 6] = rng.gamma(26.0, 1.0 / 0.15, B)                           # sig_x
    P[:, 7] = rng.gamma(15.0, 1.0 / 0.15, B)                           # sig_y
    P[:, 8] = 2.0 * rng.beta(5.0, 5.0, B) - 1.0                        # corr
    P[:, 9] = np.maximum(rng.gamma(2.0, 1.0 / 0.08, B), 2.0)           # sig_x_l
    P[:, 10] = np.maximum(rng.gamma(2.0, 1.0 / 0.14, B), 2.0)          # sig_y_l
    P[:, 11] = 2.0 * rng.beta(5.0, 5.0, B) - 1.0                       # corr_l
    P[:, 12] = rng.beta(5.0, 1.0, B)                                   # lam
    P[:, 13] = np.maximum(rng.poisson(30, B), 1)                       # n_periods
    P[:, 14] = np.clip(rng.normal(1.0, 1.0, B), 0.05, 3.0)             # mu_r
    return P


def synthetic_wind(ndays, per_day, seed=20261018):
    """AR(1) wind components with a mean drift."""
    rng = np.random.default_rng(seed)
    n = ndays * per_day
    w = np.zeros((n, 2))
    sd = 0.45
    w[0] = rng.normal(0, sd, 2)
    innov = rng.normal(0, sd * np.sqrt(1 - 0.81), (n, 2))
    for i in range(1, n):
        w[i] = 0.9 * w[i - 1] + innov[i]
    w += np.array([0.15, -0.10])
    return w.reshape(ndays, per_day, 2)


def load_workload(name, cpu=False):
    """Write the synthetic series as a wind file and read it back through
    get_wind_data (the reference's input path)."""
    if name == 'kalbar_batch512':
        return site_wind('kalbar', cpu)
    ndays, per_day, interp, rad_dist, rad_res = WORKLOADS[name]
    raw = synthetic_wind(ndays, per_day)
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, 'synthetic')
        with open(prefix + 'wind.txt', 'w') as fobj:
            for d in range(ndays):
                for wx, wy in raw[d]:
                    fobj.write('%d\t%.15g\t%.15g\n' % (d + 1, wx, wy))
        wind_data, days = _wind_reader(cpu)(prefix, interp, '00:00')
    return stack_wind(wind_data, days), wind_data, days, rad_dist, rad_res


# ---------------------------------------------------------------------------
# CPU arms on a bounded sample: the oracle port, or the unmodified reference
# ---------------------------------------------------------------------------
def _oracle_day(args):
    import warnings
    from oracle import pm_oracle as PO
    day, wind_data, rad_dist, rad_res, start_time = args
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return PO.prob_mass(day, wind_data, HPARAMS, DPARAMS, DLPARAMS, MU_R, N_PERIODS, rad_dist, rad_res, start_time)


def _reference_day(args):
    """The reference's own prob_mass (per-cell mvnun loop, ParasitoidModel.py:384-613) from baseline/_ref."""
    import warnings
    from oracle import ref_loader
    pm, _, _ = ref_loader.load()
    day, wind_data, rad_dist, rad_res, start_time = args
    with warnings.catch_warnings(), ref_loader.quiet():
        warnings.simplefilter('ignore')
        return pm.prob_mass(day, wind_data, HPARAMS, DPARAMS, DLPARAMS, MU_R, N_PERIODS, rad_dist, rad_res, start_time)


def _warm(kind):
    """Import the modules a worker needs before the clock starts (scipy.stats alone takes a second or two)."""
    from oracle import pm_oracle, cs_oracle      # noqa: F401
    if kind == 'reference':
        from oracle import ref_loader
        ref_loader.load()
    time.sleep(0.2)                              # keep this worker busy so that every worker of the pool gets one
    return 0


def _pool(n):
    import multiprocessing as mp
    # forkserver: the parent may already hold a CUDA context (get_wind_data interpolates on the device)
    return mp.get_context('forkserver').Pool(n)


def _trim(wind_data, days, n):
    return {d: wind_data[d] for d in days[:n + 1]}


def cpu_sample(wind_data, days, rad_dist, rad_res, n_kernel_days, n_chain_steps, kind='port'):
    """Time the CPU path on a bounded sample.  Phase 1: one task per day through a multiprocessing pool as
    Run.py:422-425 does.  kind 'port': `n_kernel_days` whole days of the oracle's prob_mass.  kind 'reference': the
    reference's prob_mass (54 s per day and core) on the last 1/20 of `cores` days (start_time = 0.95: 72 of the
    1440 take-off periods), scaled by 20.  Phase 2: `n_chain_steps` chain steps, single-threaded as scipy.fftpack is
    (the reference's get_solutions incl. its Python r_small_vals loop, or the oracle's).  Returns a dict."""
    from scipy import sparse
    cores = os.cpu_count() or 1
    if kind == 'reference':
        from oracle import ref_loader
        _, cs, _ = ref_loader.load()
        frac, n_kernel_days = 0.05, max(n_kernel_days, min(cores, len(days) - 1))
        pool_size = min(cores, n_kernel_days)
        sub = _trim(wind_data, days, n_kernel_days)
        with _pool(pool_size) as pool:
            pool.map(_warm, ['reference'] * pool_size, chunksize=1)
            t0 = time.perf_counter()
            pool.map(_reference_day, [(d, sub, rad_dist, rad_res, 1.0 - frac) for d in days[:n_kernel_days]], chunksize=1)
            t_k = (time.perf_counter() - t0) / frac
        # chain inputs: whole-day kernels from the port (the reference would need ~1 min per day for them)
        with _pool(min(cores, n_chain_steps + 1)) as pool:
            pmfs = pool.map(_oracle_day, [(d, sub, rad_dist, rad_res, None) for d in days[:n_chain_steps + 1]])
        get_solutions = cs.get_solutions
    else:
        from oracle import cs_oracle as CO
        pool_size = min(cores, n_kernel_days)
        sub = _trim(wind_data, days, n_kernel_days)
        with _pool(pool_size) as pool:
            pool.map(_warm, ['port'] * pool_size, chunksize=1)
            t0 = time.perf_counter()
            pmfs = pool.map(_oracle_day, [(d, sub, rad_dist, rad_res, None) for d in days[:n_kernel_days]], chunksize=1)
            t_k = time.perf_counter() - t0
        get_solutions = CO.get_solutions
    D = 2 * rad_res + 1
    ms = np.array([max(p.shape[0] for p in pmfs)] * 2)
    off = rad_res - pmfs[0].shape[0] // 2
    sol = [sparse.coo_matrix((pmfs[0].data, (pmfs[0].row + off, pmfs[0].col + off)), shape=(D, D))]
    nst = min(n_chain_steps, len(pmfs) - 1)
    t0 = time.perf_counter()
    if kind == 'reference':
        from oracle import ref_loader
        with ref_loader.quiet():
            get_solutions(sol, pmfs, days, nst + 1, D, ms)
    else:
        get_solutions(sol, pmfs, days, nst + 1, D, ms)
    t_c = time.perf_counter() - t0
    # a full pool keeps `cores` days in flight: per-day wall = one task's time / concurrency
    per_day_kernel_full_pool = (t_k * pool_size / n_kernel_days) / cores
    per_step_chain = t_c / max(nst, 1)
    return dict(kernel_s_per_day=per_day_kernel_full_pool, kernel_sample_wall_s=t_k, chain_s_per_day=per_step_chain,
                pool=pool_size, cores=cores, kernel_days=n_kernel_days, chain_steps=nst, kind=kind,
                days_per_s=1.0 / (per_day_kernel_full_pool + per_step_chain))


def cpu_sample_text(c):
    if c['kind'] == 'reference':
        return ('reference prob_mass (unmodified, mvnun shim) on the last 1/20 of %d days through a pool of %d, scaled by 20, + %d chain '
                'steps of the reference get_solutions single-threaded; per-day costs extrapolated to a pool of all %d cores'
                % (c['kernel_days'], c['pool'], c['chain_steps'], c['cores']))
    return ('%d kernel days through a pool of %d (oracle prob_mass) + %d chain steps single-threaded (oracle '
            'get_solutions), per-day costs extrapolated to a pool of all %d cores' % (c['kernel_days'], c['pool'], c['chain_steps'], c['cores']))


def cpu_full_site(site, model):
    """The oracle port on one of the 801^2 configurations, whole run (C1/C2/C3): phase 1 through a pool of all cores,
    chain single-threaded.  model: 'prob' or 'pop'.  Returns (seconds, days)."""
    from oracle import cs_oracle as CO
    from scipy import sparse
    wind, wind_data, days, rad_dist, rad_res = site_wind(site, cpu=True)
    r_dur, r_number, r_start = (1, 130000.0, None) if site == 'kalbar' else (5, 40000.0, 0.354)
    t0 = time.perf_counter()
    with _pool(min(os.cpu_count() or 1, len(days))) as pool:
        pmfs = pool.map(_oracle_day, [(d, {k: wind_data[k] for k in (d, d + 1) if k in wind_data}, rad_dist, rad_res,
                                        r_start if (model == 'pop' and i == 0) else None) for i, d in enumerate(days)])
    D = 2 * rad_res + 1
    ms = [max(p.shape[0] for p in pmfs)] * 2

    def rec(p):
        off = rad_res - p.shape[0] // 2
        return sparse.coo_matrix((p.data, (p.row + off, p.col + off)), shape=(D, D))
    if model == 'prob':
        sol = [rec(pmfs[0])]
        CO.get_solutions(sol, pmfs, days, len(days), D, ms)
    else:
        CO.get_populations([rec(p).tocsr() for p in pmfs[:r_dur]], pmfs, days, len(days), D, ms, r_dur, r_number, lambda day: 1.0 / r_dur)
    return time.perf_counter() - t0, len(days)


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                clk = float(f[1])
                smax = float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        if not sm:      # region shorter than the sampling period: use the closest sample
            for t, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(',')]
                try:
                    sm.append(float(f[1]))
                    smax = float(f[2])
                except (ValueError, IndexError):
                    pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fobj:
            return float(json.load(fobj)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def chain_bytes(P, D, flagged):
    """Algorithmic bytes of one simulated day of the chain."""
    b = 64.0 * P * P + 16.0 * D * D
    if flagged:
        b += 24.0 * P * P + 8.0 * D * D
    return b


def traffic_from_profiles(kernel):
    """dram bytes per launch of `kernel` from the committed ncu summary, if any."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    try:
        with open(path) as fobj:
            return float(json.load(fobj)[kernel]['dram_bytes_per_launch'])
    except Exception:
        return None


def phase1_roofline(prof_period, ndays, periods, steps):
    """FP64 roofline of kernel construction (k_period): executed fp64 flops per launch from the committed ncu
    capture (profiles/ncu_phase1.json: smsp__sass_thread_inst_executed_op_d{fma,mul,add}_pred_on of one C4 launch,
    60 days x 1440 periods) against the DFMA rate measured by tools/microbench.cu on this pool
    (profiles/r1_microbench.jsonl), and the exp() rate against the measured fp64 exp rate."""
    cnt, ms = prof_period
    if not cnt:
        return None
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_phase1.json')) as fobj:
            k = json.load(fobj)['k_period']
        flops = 2.0 * k['dfma'] + k['dmul'] + k['dadd']
        src = k.get('source')
    except Exception:
        return None
    peak_tf, peak_exp = 36.43, 905.0
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_microbench.jsonl')) as fobj:
            for ln in fobj:
                rec = json.loads(ln)
                if rec.get('bench') == 'dfma':
                    peak_tf = rec['tflops']
                if rec.get('bench') == 'exp_f64':
                    peak_exp = rec['gexp_per_s']
    except Exception:
        pass
    t = ms / cnt / 1000.0
    # exp() calls of the column recurrence (phase1.cuh): per period a (2h+2)^2 corner lattice, h = 23 at the default
    # covariance: 48 columns x 4 segments of 12 corners x 6 Gauss-Legendre nodes x 2 exp
    exps = float(ndays) * periods * 48 * 4 * 6 * 2
    ach = flops / t / 1e12
    return {'kernel': 'k_period', 'bound': 'fp64', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
            'executed_fp64_flops_per_launch': flops, 'avg_launch_ms': ms / cnt, 'launches': cnt,
            'exp_rate_gexp_per_s': exps / t / 1e9, 'exp_peak_gexp_per_s': peak_exp, 'exp_frac': exps / t / 1e9 / peak_exp,
            'peak_source': 'measured DFMA / exp microbenchmark (profiles/r1_microbench.jsonl)', 'flops_source': src}


def run_extras(args, ctx, torch, dist, rank, world, local, main_model, main_wind_dev, main_wind_shape, main_ndays, rad_dist, rad_res, cpu_sites):
    """The other BASELINE.json configurations, measured in the same run (see the module docstring)."""
    from parasitoids_b200 import Run, batch
    extra = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def timed(fn, steps, warmup=1):
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        tt = torch.tensor([(time.perf_counter() - t0) * 1000.0], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0]) / steps

    cells = np.random.default_rng(7).integers(0, 2 * rad_res + 1, (1024, 2)).astype(np.int32)
    # ---- weak_batch: one C4 proposal per rank through solve_batch (the path every N > 1 point of a scaling run uses)
    H = main_model[0]
    props = np.tile(np.array([H[1], H[2], H[3], H[4], H[5], H[6], *main_model[1], *main_model[2], H[0], main_model[4], main_model[3]]), (world, 1))
    props[:, 6] *= 1 + 0.01 * np.arange(world)
    props[:, 7] *= 1 - 0.005 * np.arange(world)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ms = timed(lambda: batch.solve_batch(None, props, cells, main_ndays, rad_dist, rad_res, prob_model=True, device=local,
                                             wind_device_ptr=main_wind_dev.data_ptr(), wind_shape=main_wind_shape), max(2, min(args.steps, 3)))
    extra['weak_batch'] = {'workload': args.workload, 'value': world * main_ndays / (ms / 1000.0), 'unit': 'days/s', 'ms_per_step': ms,
                           'scaling': 'weak', 'api': 'batch.solve_batch, one proposal per rank, wind resident on the device, 1024 sample cells'}
    # ---- c5: 512 prior draws x the Kalbar population solve, sharded over the ranks (strong scaling)
    wind, wind_data, days, rd, rr = site_wind('kalbar')
    wdev = torch.from_numpy(wind).cuda(local)
    wpin = torch.from_numpy(wind).pin_memory()
    proposals = prior_proposals(BATCH)
    kcells = np.random.default_rng(7).integers(0, 2 * rr + 1, (1024, 2)).astype(np.int32)
    kw = dict(prob_model=False, r_dur=1, r_number=130000.0, device=local)
    l0 = ctx.launch_count()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ms_dev = timed(lambda: batch.solve_batch(None, proposals, kcells, len(days), rd, rr, wind_device_ptr=wdev.data_ptr(),
                                                 wind_shape=wind.shape, **kw), 2)
        launches = (ctx.launch_count() - l0) // 3
        ms_e2e = timed(lambda: batch.solve_batch(wpin.numpy(), proposals, kcells, len(days), rd, rr, **kw), 2, warmup=0)
    nd5 = BATCH * len(days)
    per_rank = -(-BATCH // world)
    extra['c5'] = {'workload': 'kalbar_batch512', 'value': nd5 / (ms_dev / 1000.0), 'unit': 'days/s', 'ms_per_step': ms_dev, 'scaling': 'strong',
                   'proposals': BATCH, 'days': len(days), 'gpu_launches_per_step_rank0': int(launches),
                   'e2e': {'value': nd5 / (ms_e2e / 1000.0), 'unit': 'days/s', 'ms_per_step': ms_e2e,
                           'h2d_bytes_per_step': int(wind.nbytes + proposals[:per_rank].nbytes + kcells.nbytes),
                           'd2h_bytes_per_step': int(BATCH * len(days) * 1024 * 8)},
                   'api': 'batch.solve_batch: proposals sharded over the ranks, one all_gather of the sampled cells'}
    del wdev
    # ---- c1 / c2 / c3: the 801^2 configurations through Run.solve, COO triplets of every day to the host (N = 1 only)
    if world == 1:
        for name, site, model in (('c1', 'kalbar', 'prob'), ('c2', 'kalbar', 'pop'), ('c3_prob', 'carnarvon', 'prob'), ('c3_pop', 'carnarvon', 'pop')):
            wind, wind_data, days, rd, rr = site_wind(site)
            r_dur, r_number, r_start = (1, 130000.0, None) if site == 'kalbar' else (5, 40000.0, 0.354)
            skw = dict(prob_model=True) if model == 'prob' else dict(prob_model=False, r_dur=r_dur, r_number=r_number,
                                                                     r_dist=[1.0 / r_dur] * r_dur, r_start=r_start)
            wp = torch.from_numpy(wind).pin_memory().numpy()
            m = (HPARAMS, DPARAMS, DLPARAMS, MU_R, N_PERIODS, rd, rr)

            def one():
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    res = Run.solve(wp, len(days), *m, want_coo='csr', device=local, **skw)
                    res.csr_arrays()
                    res.close()
            ms = timed(one, 5, warmup=2)
            rec = {'workload': '%s %s model, 801x801, %d days (Run.py presets)' % (site, 'probability' if model == 'prob' else 'population', len(days)),
                   'e2e': {'value': len(days) / (ms / 1000.0), 'unit': 'days/s', 'ms_per_step': ms,
                           'api': "Run.solve(want_coo='csr'): wind from pinned host memory, CSR arrays of all days to host"}}
            if name in cpu_sites:
                rec['cpu_baseline'] = cpu_sites[name]
            extra[name] = rec
    return extra


def run_single_solve(args, ctx, torch, dist, rank, world, local, wind, ndays, rad_dist, rad_res):
    """--single-solve: one forward solve over all ranks; value = days of that ONE solve per second (strong scaling)."""
    from parasitoids_b200 import multi
    model = (HPARAMS, DPARAMS, DLPARAMS, MU_R, N_PERIODS, rad_dist, rad_res)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one():
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return multi.solve_single(wind, ndays, *model, device=local)
    res = None
    for _ in range(max(args.warmup, 1)):
        res = one()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.25)
    l0 = ctx.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        res = one()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    clk = clocks.stop(t0, t1)
    tt = torch.tensor([e0.elapsed_time(e1), (t1 - t0) * 1000.0], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, wall_ms = float(tt[0]), float(tt[1])
    launches = ctx.launch_count() - l0
    # a checksum every rank can produce from its own rows: the mass of the last day (summed over the ranks)
    mass = res.rows[-1].sum().reshape(1)
    if world > 1:
        dist.all_reduce(mass)
    if rank == 0:
        D, per = res.dom_len, int(res.rows.shape[1])
        Nc = res.N // 2 + 1
        cg = -(-(-(-Nc // world)) // 4) * 4
        xchg = 16.0 * cg * per * (world - 1)            # bytes one GPU sends per day (its columns x the other ranks' rows)
        line = {'metric': 'simulated days/sec (fp64)', 'value': ndays * args.steps / (t_ms / 1000.0), 'unit': 'days/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': max(args.warmup, 1), 'ms_per_step': t_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': args.workload, 'days': ndays, 'dom_len': D, 'torus_P': res.P, 'fft_len': res.N,
                           'parallelism': ('ONE solve over %d GPU(s): phase-1 days round-robin + one all-gather of the kernels; chain state '
                                           'sharded by spectral column, one all-to-all and one 4-double all-gather per day' % world),
                           'sharded': bool(res.sharded), 'rows_per_rank': per,
                           'l2': 'per-rank working set per day exceeds the 126 MB L2 up to 4 GPUs; no explicit flush'},
                'wall_ms_per_step': wall_ms / args.steps, 'clocks': clk, 'gpu_launches': int(launches),
                'timer': 'CUDA events on the stream every kernel and collective of the solve is enqueued on, max over ranks; wind uploaded from host memory inside the timed region',
                'nvlink': {'alltoall_bytes_sent_per_gpu_per_day': xchg, 'bytes_per_solve_per_gpu': xchg * (ndays - 1),
                           'time_at_770_GBps_ms_per_solve': xchg * (ndays - 1) / 770e9 * 1e3},
                'last_day_mass': float(mass[0]), 'max_outside_domain': float(res.meta[:, 3].max()),
                'e2e': {'value': ndays * args.steps / (wall_ms / 1000.0), 'unit': 'days/s', 'h2d_bytes_per_step': int(wind.nbytes),
                        'd2h_bytes_per_step': int(res.meta.nbytes), 'note': 'solutions stay sharded on the GPUs (each rank holds its rows of every day)'}}
        print(json.dumps(line))


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return
    wind, wind_data, days, rad_dist, rad_res = load_workload(args.workload, cpu=True)
    from oracle import ref_loader
    kind = 'reference' if ref_loader.available() else 'port'
    small = rad_res < 1000
    total = args.warmup + args.steps
    nk, nc = (8, 4) if small else (4, 3 if total <= 3 else 2)
    vals, last = [], None
    for i in range(total):
        last = cpu_sample(wind_data, days, rad_dist, rad_res, nk, nc, kind)
        if i >= args.warmup:
            vals.append(last['days_per_s'])
    v = float(np.mean(vals))
    line = {'impl': 'reference', 'metric': 'simulated days/sec (fp64)', 'value': v, 'unit': 'days/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * WORKLOADS[args.workload][0] / v,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': args.workload,
                       'note': ('CPU path: the unmodified reference modules (baseline/_ref, scipy.stats.mvn.mvnun shimmed onto SciPy\'s Genz BVU)'
                                if kind == 'reference' else 'CPU path: numpy/scipy restatement of the reference (oracle/)') + ', host cores only'},
            'cpu_baseline': {'value': v, 'unit': 'days/s', 'cores': last['cores'], 'kind': kind, 'sample': cpu_sample_text(last),
                             'kernel_s_per_day': last['kernel_s_per_day'], 'chain_s_per_day': last['chain_s_per_day']},
            'e2e': {'value': v, 'unit': 'days/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='synthetic_4097x4097_60d', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the extra configurations (c1-c3, c5, weak_batch)')
    ap.add_argument('--single-solve', action='store_true',
                    help='ONE solve of the workload sharded over all ranks (parasitoids_b200.multi: phase-1 days round-robin, '
                         'slab-decomposed spectral chain with one all-to-all per day) -- strong scaling of BASELINE config 4')
    ap.add_argument('--opt', action='append', default=[], metavar='KEY=VALUE',
                    help='library option (pkb_set_option), e.g. fuse_rows=0; recorded in config')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.gpus > 1 and 'WORLD_SIZE' not in os.environ:
        # launched bare: re-launch as one process per GPU (the driver does this itself)
        os.execvp(sys.executable, [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
                                   '--master-addr', '127.0.0.1', '--master-port', os.environ.get('MASTER_PORT', '29541'),
                                   os.path.abspath(__file__)] + sys.argv[1:])
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    ndays = WORKLOADS[args.workload][0]

    # CPU baseline first (rank 0 of a one-GPU run only): the oracle port on a bounded sample of the same workload
    cpu, cpu_sites = None, {}
    if world == 1 and not args.no_cpu_baseline:
        _, wd_cpu, days_cpu, rd_cpu, rr_cpu = load_workload(args.workload, cpu=True)
        nk, nc = (8, 4) if rr_cpu < 1000 else (4, 3)
        c = cpu_sample(wd_cpu, days_cpu, rd_cpu, rr_cpu, nk, nc)
        cpu = {'value': c['days_per_s'], 'unit': 'days/s', 'cores': c['cores'], 'kind': 'port', 'sample': cpu_sample_text(c),
               'kernel_s_per_day': c['kernel_s_per_day'], 'chain_s_per_day': c['chain_s_per_day']}
        if args.workload == 'synthetic_4097x4097_60d' and not args.no_extras:
            for name, site, model in (('c1', 'kalbar', 'prob'), ('c2', 'kalbar', 'pop'), ('c3_prob', 'carnarvon', 'prob'), ('c3_pop', 'carnarvon', 'pop')):
                sec, nd_site = cpu_full_site(site, model)
                cpu_sites[name] = {'value': nd_site / sec, 'unit': 'days/s', 'seconds': sec, 'cores': c['cores'], 'kind': 'port',
                                   'sample': 'whole run: oracle prob_mass through a pool of all cores, oracle chain single-threaded'}
        del wd_cpu

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    if world > 1:
        # With NCCL_DEBUG >= VERSION (this image's default) NCCL printf()s "NCCL version ..." to stdout when the
        # first communicator is created -- next to the one JSON line the driver parses.  File descriptor 1 points
        # at stderr while that happens (init + one barrier), then it is restored.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    from parasitoids_b200 import Run, _lib, batch
    ctx = _lib.ctx(local)
    for kv in args.opt:
        key, val = kv.split('=')
        ctx.set_option(key, float(val))
    wind, wind_data, days, rad_dist, rad_res = load_workload(args.workload)
    if args.single_solve:
        run_single_solve(args, ctx, torch, dist, rank, world, local, wind, ndays, rad_dist, rad_res)
        if world > 1:
            dist.destroy_process_group()
        return

    # N > 1: a likelihood batch of one parameter proposal per rank (proposal 0 = the defaults),
    # sharded by parasitoids_b200.batch.solve_batch; the only collective is its all_gather
    model = (HPARAMS, DPARAMS, DLPARAMS, MU_R, N_PERIODS, rad_dist, rad_res)
    batch_mode = args.workload == 'kalbar_batch512'
    proposals = np.tile(np.array([HPARAMS[1], HPARAMS[2], HPARAMS[3], HPARAMS[4], HPARAMS[5], HPARAMS[6], *DPARAMS, *DLPARAMS,
                                  HPARAMS[0], N_PERIODS, MU_R]), (world, 1))
    proposals[:, 6] *= 1 + 0.01 * np.arange(world)
    proposals[:, 7] *= 1 - 0.005 * np.arange(world)
    pop_kw = dict(prob_model=True)
    if batch_mode:
        proposals = prior_proposals(BATCH)
        pop_kw = dict(prob_model=False, r_dur=1, r_number=130000.0)      # Run.py:126-138 (kalbar preset)
    units = (BATCH if batch_mode else world) * ndays                     # simulated days per step, all ranks
    wind_dev = torch.from_numpy(wind).cuda(local)
    wind_pinned = torch.from_numpy(wind).pin_memory()
    rng = np.random.default_rng(7)
    cells = rng.integers(0, 2 * rad_res + 1, (1024, 2)).astype(np.int32)

    class _Info(object):
        pass

    def step_device():
        if batch_mode:
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')      # proposals with strong drift warn about wasps leaving the domain
                batch.solve_batch(None, proposals, cells, ndays, rad_dist, rad_res, device=local,
                                  wind_device_ptr=wind_dev.data_ptr(), wind_shape=wind.shape, **pop_kw)
            return None
        return Run.solve(None, ndays, *model, want_coo=False, keep_device=True, wind_device_ptr=wind_dev.data_ptr(),
                         wind_shape=wind.shape, device=local)

    def step_e2e():
        if batch_mode:
            # host wind in, sampled cells of every proposal out (what the likelihood consumes)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                out = batch.solve_batch(wind_pinned.numpy(), proposals, cells, ndays, rad_dist, rad_res, device=local, **pop_kw)
            return None, out.size // 2        # counted below as 16 bytes per entry
        res = Run.solve(wind_pinned.numpy(), ndays, *model, want_coo='csr', device=local)
        off, rowoff, cols, vals = res.csr_arrays()
        return res, int(off[-1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    # ---- device-resident measurement -------------------------------------------
    def close(r):
        if r is not None:
            r.close()

    # geometry of the workload (one untimed solve on every rank)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        r = Run.solve(None, ndays, *model, want_coo=False, keep_device=True, wind_device_ptr=wind_dev.data_ptr(),
                      wind_shape=wind.shape, device=local)
    info = (r.P, r.N, r.dom_len, r.flags(), r.radii())
    window_steps = r.window_steps()
    spectral_steps = len(r.spectral_steps())
    r.close()
    for _ in range(args.warmup):
        close(step_device())
    barrier()

    def timed_pass(profile):
        """K steps between two events on the library's stream; returns (device ms, wall ms, launches, clocks, phase ms)."""
        if profile:
            ctx.profile_reset()
            ctx.profile(True)
        clocks = ClockSampler(local)
        clocks.start()
        time.sleep(0.25)
        l0 = ctx.launch_count()
        barrier()
        t0 = time.perf_counter()
        ctx.mark(0)
        chain = 0.0
        for _ in range(args.steps):
            close(step_device())
            chain += ctx.timing()['chain_ms']      # phase-2 time of the LAST solve of the step
        ctx.mark(1)
        barrier()
        t1 = time.perf_counter()
        dev = ctx.elapsed_ms(0, 1)
        clk = clocks.stop(t0, t1)
        n = ctx.launch_count() - l0
        if profile:
            ctx.profile(False)
        return max(dev, 0.0), (t1 - t0) * 1000.0, n, clk, chain

    # pass A: the reported value (no per-kernel events); pass B: the same K steps with every
    # launch bracketed by CUDA events on its stream -> per-kernel durations for the roofline
    dev_ms, wall_ms, launches, clk, chain_total_ms = timed_pass(False)
    prof_ms, _, _, _, _ = timed_pass(True)
    tt = torch.tensor([dev_ms, wall_ms, prof_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, wall_ms, prof_ms = float(tt[0]), float(tt[1]), float(tt[2])
    value = units * args.steps / (t_ms / 1000.0)

    # ---- per-kernel roofline (rank 0) --------------------------------------------
    P, N, D, flags, radii = info
    kernels = ['k_rows_fwd', 'k_cols', 'k_rows_inv', 'k_rows_fwd_win', 'k_cols_win', 'k_rows_inv_win', 'k_kernel_rows_win',
               'k_kernel_rows', 'k_kernel_rows_batch', 'k_emit_dense', 'k_step_finalize',
               'k_zero_pad', 'k_period', 'k_day_finalize', 'k_drift', 'k_hprob', 'k_bvn_setup', 'k_copy_domain', 'k_place_kernel',
               'k_stencil', 'k_row_stats']
    prof = {k: ctx.profile_get(k) for k in kernels}
    chain_k = ['k_rows_fwd', 'k_cols', 'k_rows_inv', 'k_kernel_rows', 'k_kernel_rows_batch', 'k_emit_dense', 'k_step_finalize', 'k_zero_pad']
    phase1_ms = sum(prof[k][1] for k in ('k_period', 'k_day_finalize', 'k_drift', 'k_hprob', 'k_bvn_setup'))
    nsteps_chain = prof['k_cols'][0] + prof['k_cols_win'][0]
    peak, peak_src = peaks()
    # algorithmic bytes per launch of each chain kernel = its share of B_day.  Support-window
    # steps run the same three kernels on a torus sized for the (numerical) support of the state; they are profiled
    # under k_*_win and belong to the same family: a launch of either kind advances one simulated day.
    share = {'k_rows_fwd': 8.0 * P * P,                      # kernel/state row pass: write one half-spectrum
             'k_cols': 40.0 * P * P,                         # column pass + multiply (24 P^2) and inverse column pass (16 P^2)
             'k_rows_inv': 16.0 * P * P,                     # inverse row pass: read half-spectrum, write real grid
             'k_emit_dense': 16.0 * D * D}                   # renormalised output: read + write the domain
    fam = {k: (prof[k][0] + prof.get(k + '_win', (0, 0.0))[0], prof[k][1] + prof.get(k + '_win', (0, 0.0))[1]) for k in share}
    dom = max(('k_rows_fwd', 'k_cols', 'k_rows_inv'), key=lambda k: fam[k][1])
    cnt, ms = fam[dom]
    roofline = None
    if cnt:
        ach = share[dom] / (ms / cnt / 1000.0) / 1e9
        roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'traffic': None if batch_mode else traffic_from_profiles(dom), 'peak_source': peak_src, 'launches': cnt,
                    'avg_launch_ms': ms / cnt, 'algorithmic_bytes_per_launch': share[dom],
                    'launches_on_support_windows': prof.get(dom + '_win', (0, 0.0))[0],
                    'note': 'algorithmic bytes are SURVEY.md 8d\'s share of B_day(P) for one simulated day; launches on a support window move '
                            'fewer bytes than that (cells below 1e-15 are neither transformed nor stored), which is how frac can exceed 1',
                    'family_ms_per_solve': {k: round(fam[k][1] / args.steps, 4) for k in fam}}
        if prof[dom][0]:
            # the same kernel on its whole-torus launches only (the geometry the algorithmic bytes describe; in the fused C4 solve
            # those are the spectral-resident steps of the last days, whose k_cols starts from the stored spectrum)
            wms = prof[dom][1] / prof[dom][0]
            roofline['whole_torus_launches'] = {'launches': prof[dom][0], 'avg_launch_ms': wms, 'achieved': share[dom] / (wms / 1000.0) / 1e9,
                                                'frac': share[dom] / (wms / 1000.0) / 1e9 / peak}
    roofline_chain = None
    if nsteps_chain and not batch_mode and chain_total_ms > 0:      # batch mode: proposals differ in torus size; the per-kernel roofline above still applies
        nflag = sum(1 for f in flags if f)
        bytes_solve = sum(chain_bytes(P, D, f) for f in flags[1:])
        # whole chain phase of the solve (library events around phase 2, pass A): kernels, gaps and overlap included
        ach = bytes_solve * args.steps / (chain_total_ms / 1000.0) / 1e9
        roofline_chain = {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                          'algorithmic_bytes_per_day': chain_bytes(P, D, False), 'flagged_days': nflag,
                          'chain_ms_per_day': chain_total_ms / args.steps / max(ndays - 1, 1),
                          'phase1_kernel_ms_per_day': phase1_ms / (ndays * args.steps),
                          'profiled_pass_ms_per_step': prof_ms / args.steps,
                          'kernel_ms': {k: round(prof[k][1] / args.steps, 4) for k in kernels if prof[k][0]}}

    # ---- end to end through the public API, host buffers -------------------------
    for _ in range(max(1, min(args.warmup, 2))):
        r, _ = step_e2e()
        close(r)
    barrier()
    ctx.mark(2)
    te0 = time.perf_counter()
    nnz_tot = 0
    for _ in range(args.steps):
        r, nnz = step_e2e()
        nnz_tot += nnz
        close(r)
    ctx.mark(3)
    barrier()
    e_wall = (time.perf_counter() - te0) * 1000.0
    te = torch.tensor([e_wall], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_ms = float(te[0])
    e2e = {'value': units * args.steps / (e_ms / 1000.0), 'unit': 'days/s',
           # batch modes: every rank uploads the wind once per call plus its share of the proposals and the cells
           # (batch mode: every rank uploads the wind once per call plus its share of the proposals and the cells; otherwise one solve per rank)
           'h2d_bytes_per_step': int(wind.nbytes + proposals[:-(-len(proposals) // world)].nbytes + cells.nbytes) if batch_mode else int(world * wind.nbytes),
           'd2h_bytes_per_step': (int(nnz_tot / args.steps * 16 + (ndays + 1) * 8) if batch_mode else
                                  int(world * (nnz_tot / args.steps * 12 + ndays * D * 8 + (ndays + 1) * 8))),
           'ms_per_step': e_ms / args.steps, 'timer': 'host wall clock between device synchronisations',
           'api': ('parasitoids_b200.batch.solve_batch: wind from pinned host memory on every rank, sampled cells of all proposals '
                   'all-gathered and copied to host') if batch_mode else
                  "parasitoids_b200.Run.solve(want_coo='csr'): wind from pinned host memory; the thresholded, renormalised solution of every day to "
                  'host as CSR arrays (row offsets, int32 column, fp64 value: what Run.main saves, Run.py:490-510)'}

    extra = None
    if not args.no_extras and not batch_mode and args.workload == 'synthetic_4097x4097_60d':
        extra = run_extras(args, ctx, torch, dist, rank, world, local, model, wind_dev, wind.shape, ndays, rad_dist, rad_res, cpu_sites)
    if rank == 0:
        line = {'metric': 'simulated days/sec (fp64)', 'value': value, 'unit': 'days/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': t_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong' if batch_mode else 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': args.workload, 'days': ndays, 'dom_len': D, 'torus_P': P, 'fft_len': N,
                           'periods_per_day': int(wind.shape[1]), 'kernel_radius_min_max': [int(min(radii)), int(max(radii))],
                           'support_window_steps': window_steps, 'spectral_resident_steps': spectral_steps,
                           'parallelism': ('likelihood batch of %d proposals sharded over %d GPU(s) (batch.solve_batch), one all_gather of 1024 sampled cells x days per step' % (len(proposals), world)) if batch_mode
                                          else ('single solve' if world == 1 else 'one independent solve per GPU (%d replicas of the N = 1 step, max over ranks, no collective)' % world),
                           'l2': 'every solve writes %.1f GB of dense daily solutions (> 126 MB L2) between two timed solves; whole-torus chain '
                                 'steps work on %.0f MB; no explicit flush' % (ndays * 8.0 * D * D / 1e9, 3 * 8.0 * P * P / 1e6)},
                'wall_ms_per_step': wall_ms / args.steps, 'clocks': clk, 'e2e': e2e, 'gpu_launches': int(launches),
                'roofline': roofline, 'roofline_chain': roofline_chain,
                'roofline_phase1': None if batch_mode else phase1_roofline(prof['k_period'], ndays, int(wind.shape[1]), args.steps)}
        if extra:
            line['extra'] = extra
        if args.opt:
            line['config']['options'] = list(args.opt)
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
