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
            thresholded COO triplets of every day copied back inside the timed
            region -- what a caller of Run.main sees.
`roofline`  the dominant chain kernel against the measured HBM copy bandwidth;
`roofline_chain` all chain kernels of one simulated day against B_day(P, D).
`cpu_baseline` / `--impl reference`: the CPU path (oracle/, the numpy
            restatement of the reference -- the reference itself is Python
            that cannot travel to the GPU box) on a bounded sample, phase 1
            fanned out over a multiprocessing pool like Run.py:422-425.

N > 1: independent parameter proposals of the same solve, one per rank (the
batched-likelihood partitioning of SURVEY.md section 8e); the only collective
is one NCCL all_gather of the sampled-cell outputs per step.
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


def kalbar_wind():
    """Kalbar wind series (data/kalbarwind.txt as committed in tests/golden/wind.npz) through get_wind_data."""
    from parasitoids_b200 import ParasitoidModel as PM
    from parasitoids_b200 import Run
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wind.npz'))
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, 'kalbar')
        with open(prefix + 'wind.txt', 'w') as fobj:
            for d, block in zip(z['kalbar_days'], z['kalbar_raw']):
                for wx, wy, _ in block:
                    fobj.write('%d\t%.17g\t%.17g\n' % (d, wx, wy))
        wind_data, days = PM.get_wind_data(prefix, 30, '00:00')
    return Run.stack_wind(wind_data, days), wind_data, days, 10000.0, 400


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


def load_workload(name):
    """Write the synthetic series as a wind file and read it back through the
    package's own get_wind_data (the reference's input path)."""
    from parasitoids_b200 import ParasitoidModel as PM
    from parasitoids_b200 import Run
    if name == 'kalbar_batch512':
        return kalbar_wind()
    ndays, per_day, interp, rad_dist, rad_res = WORKLOADS[name]
    raw = synthetic_wind(ndays, per_day)
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, 'synthetic')
        with open(prefix + 'wind.txt', 'w') as fobj:
            for d in range(ndays):
                for wx, wy in raw[d]:
                    fobj.write('%d\t%.15g\t%.15g\n' % (d + 1, wx, wy))
        wind_data, days = PM.get_wind_data(prefix, interp, '00:00')
    return Run.stack_wind(wind_data, days), wind_data, days, rad_dist, rad_res


# ---------------------------------------------------------------------------
# CPU path (oracle) on a bounded sample
# ---------------------------------------------------------------------------
def _oracle_day(args):
    import warnings
    from oracle import pm_oracle as PO
    day, wind_data, rad_dist, rad_res = args
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return PO.prob_mass(day, wind_data, HPARAMS, DPARAMS, DLPARAMS, MU_R, N_PERIODS, rad_dist, rad_res)


def cpu_sample(wind_data, days, rad_dist, rad_res, n_kernel_days, n_chain_steps):
    """Time the CPU path on `n_kernel_days` kernels (multiprocessing pool, one
    task per day as Run.py:422-425) and `n_chain_steps` chain steps
    (single-threaded pocketfft, as scipy.fftpack is).  Returns a dict."""
    import multiprocessing as mp
    from scipy import sparse
    from oracle import cs_oracle as CO
    cores = os.cpu_count() or 1
    pool_size = min(cores, n_kernel_days)
    sub = {d: wind_data[d] for d in days[:n_kernel_days + 1]}
    t0 = time.perf_counter()
    with mp.get_context('fork').Pool(pool_size) as pool:
        pmfs = pool.map(_oracle_day, [(d, sub, rad_dist, rad_res) for d in days[:n_kernel_days]])
    t_k = time.perf_counter() - t0
    D = 2 * rad_res + 1
    ms = [max(p.shape[0] for p in pmfs)] * 2
    off = rad_res - pmfs[0].shape[0] // 2
    sol = [sparse.coo_matrix((pmfs[0].data, (pmfs[0].row + off, pmfs[0].col + off)), shape=(D, D))]
    nst = min(n_chain_steps, len(pmfs) - 1)
    t0 = time.perf_counter()
    CO.get_solutions(sol, pmfs, days, nst + 1, D, ms)
    t_c = time.perf_counter() - t0
    # a full pool keeps `cores` days in flight: per-day wall = one task's time / concurrency
    per_day_kernel = t_k / n_kernel_days if pool_size >= n_kernel_days else t_k / n_kernel_days
    per_day_kernel_full_pool = (t_k * pool_size / n_kernel_days) / cores
    per_step_chain = t_c / max(nst, 1)
    return dict(kernel_s_per_day=per_day_kernel_full_pool, kernel_sample_wall_s=t_k, chain_s_per_day=per_step_chain,
                pool=pool_size, cores=cores, kernel_days=n_kernel_days, chain_steps=nst,
                days_per_s=1.0 / (per_day_kernel_full_pool + per_step_chain))


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


def run_reference(args, rank, world):
    """--impl reference: the CPU path on this box's host cores."""
    if rank != 0:
        return
    wind, wind_data, days, rad_dist, rad_res = load_workload(args.workload)
    small = rad_res < 1000
    nk, nc = (8, 4) if small else (4, 2)
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_sample(wind_data, days, rad_dist, rad_res, nk, nc)
        if i >= args.warmup:
            vals.append(last['days_per_s'])
    v = float(np.mean(vals))
    sample = ('%d kernel days through a pool of %d (oracle prob_mass) + %d chain steps single-threaded (oracle '
              'get_solutions), per-day costs extrapolated to a pool of all %d cores' % (nk, last['pool'], last['chain_steps'], last['cores']))
    line = {'impl': 'reference', 'metric': 'simulated days/sec (fp64)', 'value': v, 'unit': 'days/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * WORKLOADS[args.workload][0] / v,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': args.workload, 'note': 'CPU path: numpy/scipy restatement of the reference (oracle/), host cores only'},
            'cpu_baseline': {'value': v, 'unit': 'days/s', 'cores': last['cores'], 'kind': 'port', 'sample': sample,
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
    wind, wind_data, days, rad_dist, rad_res = load_workload(args.workload)

    # CPU baseline first (fork-based pool before any CUDA context exists)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        small = rad_res < 1000
        nk, nc = (8, 4) if small else (4, 2)
        c = cpu_sample(wind_data, days, rad_dist, rad_res, nk, nc)
        cpu = {'value': c['days_per_s'], 'unit': 'days/s', 'cores': c['cores'], 'kind': 'port',
               'sample': '%d kernel days through a pool of %d (oracle prob_mass) + %d chain steps single-threaded (oracle '
                         'get_solutions), per-day costs extrapolated to a pool of all %d cores'
                         % (c['kernel_days'], c['pool'], c['chain_steps'], c['cores']),
               'kernel_s_per_day': c['kernel_s_per_day'], 'chain_s_per_day': c['chain_s_per_day']}

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
        if world > 1 or batch_mode:
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')      # proposals with strong drift warn about wasps leaving the domain
                batch.solve_batch(None, proposals, cells, ndays, rad_dist, rad_res, device=local,
                                  wind_device_ptr=wind_dev.data_ptr(), wind_shape=wind.shape, **pop_kw)
            return None
        return Run.solve(None, ndays, *model, want_coo=False, keep_device=True, wind_device_ptr=wind_dev.data_ptr(),
                         wind_shape=wind.shape, device=local)

    def step_e2e():
        if world > 1 or batch_mode:
            # host wind in, sampled cells of every proposal out (what the likelihood consumes)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                out = batch.solve_batch(wind_pinned.numpy(), proposals, cells, ndays, rad_dist, rad_res, device=local, **pop_kw)
            return None, out.size // 2        # counted below as 16 bytes per entry
        res = Run.solve(wind_pinned.numpy(), ndays, *model, want_coo=True, device=local)
        off, rows, cols, vals = res.coo_arrays()
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
    nsteps_chain = prof['k_cols'][0]
    peak, peak_src = peaks()
    dom = max(chain_k, key=lambda k: prof[k][1])
    # algorithmic bytes per launch of each chain kernel = its share of B_day; the
    # per-kernel roofline is taken on the whole-torus launches only (support-window steps run the same
    # kernels on a smaller torus and are accounted separately as k_*_win)
    share = {'k_rows_fwd': 8.0 * P * P,                      # kernel/state row pass: write one half-spectrum
             'k_cols': 40.0 * P * P,                         # column pass + multiply (24 P^2) and inverse column pass (16 P^2)
             'k_rows_inv': 16.0 * P * P,                     # inverse row pass: read half-spectrum, write real grid
             'k_emit_dense': 16.0 * D * D,                   # renormalised output: read + write the domain
             'k_kernel_rows': 0.0, 'k_kernel_rows_batch': 0.0, 'k_step_finalize': 0.0, 'k_zero_pad': 0.0}
    cnt, ms = prof[dom]
    roofline = None
    if cnt:
        ach = share[dom] / (ms / cnt / 1000.0) / 1e9
        roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'traffic': None if batch_mode else traffic_from_profiles(dom), 'peak_source': peak_src, 'launches': cnt,
                    'avg_launch_ms': ms / cnt, 'algorithmic_bytes_per_launch': share[dom]}
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
           'h2d_bytes_per_step': int(wind.nbytes) * (len(proposals) if (world > 1 or batch_mode) else 1), 'd2h_bytes_per_step': int(nnz_tot / args.steps * 16 + (ndays + 1) * 8),
           'ms_per_step': e_ms / args.steps, 'timer': 'host wall clock between device synchronisations',
           'api': ('parasitoids_b200.batch.solve_batch: wind from pinned host memory on every rank, sampled cells of all proposals '
                   'all-gathered and copied to host') if (world > 1 or batch_mode) else
                  'parasitoids_b200.Run.solve(want_coo=True): wind from pinned host memory, COO triplets of all days to host'}

    if rank == 0:
        line = {'metric': 'simulated days/sec (fp64)', 'value': value, 'unit': 'days/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': t_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong' if batch_mode else 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': args.workload, 'days': ndays, 'dom_len': D, 'torus_P': P, 'fft_len': N,
                           'periods_per_day': int(wind.shape[1]), 'kernel_radius_min_max': [int(min(radii)), int(max(radii))],
                           'support_window_steps': window_steps,
                           'parallelism': ('likelihood batch of %d proposals sharded over %d GPU(s) (batch.solve_batch), one all_gather of 1024 sampled cells x days per step' % (len(proposals), world)) if (world > 1 or batch_mode) else 'single solve',
                           'l2': 'working set per chain step (%.0f MB) exceeds the 126 MB L2; no explicit flush' % (3 * 8.0 * P * P / 1e6)},
                'wall_ms_per_step': wall_ms / args.steps, 'clocks': clk, 'e2e': e2e, 'gpu_launches': int(launches),
                'roofline': roofline, 'roofline_chain': roofline_chain}
        if args.opt:
            line['config']['options'] = list(args.opt)
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
