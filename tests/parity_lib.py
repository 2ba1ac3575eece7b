"""Full-size parity measurements: the CUDA path (through the C ABI) against the
oracle on the configurations BASELINE.json names, with the NUMBERS kept.

Used by tests/test_parity_full.py (asserts the north-star bars) and by
tools/parity_report.py (writes profiles/parity_r02.json).  Everything here is
test infrastructure; the oracle is only ever the checker.

Per day the record holds (SURVEY.md section 8d "parity checks to report"):
  max_abs, rel_l1   dense PRE-threshold grids (bars 1e-10 / 1e-9)
  flips             post-threshold support symmetric difference (cells within
                    1e-4 relative of the 1e-8 drop threshold)
  mass_err          |sum - 1| of the thresholded probability day (bar 1e-12)
plus flag-sequence and kernel-radius equality for the whole solve.
"""
import json
import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import helpers as H        # noqa: E402

NEGVAL = 1e-8
OUT_JSON = os.path.join(ROOT, 'gpurun_out', 'parity_r02.json')


# ---- oracle workers (importable for a forkserver pool) ------------------------------
def oracle_day(args):
    from oracle import pm_oracle as PO
    day, wind_sub, hp, dp, dl, mu_r, n_periods, rad_dist, rad_res, start_time = args
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return PO.prob_mass(day, wind_sub, hp, dp, dl, mu_r, n_periods, rad_dist, rad_res, start_time)


def oracle_pre_window(args):
    """Oracle prob_mass of one day: the dense pre-threshold grid cropped to its support window."""
    from oracle import pm_oracle as PO
    day, sub, _, hp, dp, dl, mu_r, n_periods, rad_dist, rad_res = args
    det = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        PO.prob_mass(day, sub, hp, dp, dl, mu_r, n_periods, rad_dist, rad_res, None, details=det)
    pre = det['pmf_pre']
    I, J = np.nonzero(pre)
    r = int(max(np.abs(I - rad_res).max(), np.abs(J - rad_res).max()))
    return pre[rad_res - r:rad_res + r + 1, rad_res - r:rad_res + r + 1].copy()


def oracle_kernels(wind_data, days, model, start_time0=None, pool=None):
    """Oracle prob_mass for every day (fan-out like Run.py:422-425).  `wind_data` is trimmed per task to the
    two days a kernel reads so that the pickles stay small."""
    hp, dp, dl, mu_r, n_periods, rad_dist, rad_res = model
    tasks = []
    for i, d in enumerate(days):
        sub = {k: wind_data[k] for k in (d, d + 1) if k in wind_data}
        tasks.append((d, sub, hp, dp, dl, mu_r, n_periods, rad_dist, rad_res, start_time0 if i == 0 else None))
    if pool is None:
        return [oracle_day(t) for t in tasks]
    return pool.map(oracle_day, tasks, chunksize=1)


def oracle_proposal(args):
    """Full oracle population solve of one MCMC proposal, sampled at `cells` (Bayes_Run.py:236-306 without
    the sprd_factor branch)."""
    from oracle import cs_oracle as CO
    prop, wind_data, days, rad_dist, rad_res, r_dur, r_number, r_start, cells = args
    from parasitoids_b200 import batch
    hp, dp, dl, mu_r, n_periods = batch.unpack_proposal(prop)
    pmfs = oracle_kernels(wind_data, days, (hp, dp, dl, mu_r, n_periods, rad_dist, rad_res), r_start)
    D = 2 * rad_res + 1
    ms = [max(p.shape[0] for p in pmfs)] * 2
    r_spread = [H.recentre(p, rad_res).tocsr() for p in pmfs[:r_dur]]
    det = {}
    pop = CO.get_populations(r_spread, pmfs, days, len(days), D, ms, r_dur, r_number, lambda day: 1.0 / r_dur, details=det)
    thr = np.array([[pop[d][r, c] for r, c in cells] for d in range(len(days))])
    pre = np.array([[det['pre'][d][r, c] for r, c in cells] for d in range(len(days))])
    return thr, pre, [int(p.shape[0]) for p in pmfs], [bool(f) for f in det['flags']]


def make_pool(n=None):
    # forkserver: the parent may already hold a CUDA context
    return mp.get_context('forkserver').Pool(n or os.cpu_count() or 1)


# ---- metrics ------------------------------------------------------------------------
def day_record(got_pre, ref_pre, got_thr=None, ref_thr=None, prob=True):
    got_pre = np.asarray(got_pre, dtype=float)
    ref_pre = np.asarray(ref_pre, dtype=float)
    rec = {'max_abs': float(np.abs(got_pre - ref_pre).max()), 'rel_l1': float(H.rel_l1(got_pre, ref_pre)),
           'ref_max': float(np.abs(ref_pre).max())}
    if got_thr is not None:
        got_thr = np.asarray(got_thr, dtype=float)
        ref_thr = np.asarray(ref_thr, dtype=float)
        only = (got_thr != 0) != (ref_thr != 0)
        rec['flips'] = int(only.sum())
        if only.any():
            vals = np.where(got_thr != 0, got_thr, ref_thr)[only]
            rec['flip_max_rel_dist'] = float(np.abs(vals / NEGVAL - 1).max()) if prob else None
        both = ~only
        rec['thr_max_abs'] = float(np.abs(np.where(both, got_thr - ref_thr, 0.0)).max())
        if prob:
            rec['mass_err'] = float(abs(np.sum(got_thr, dtype=np.longdouble) - 1))
    return rec


def summarise(days):
    keys = ('max_abs', 'rel_l1', 'thr_max_abs', 'mass_err')
    out = {k: max((d[k] for d in days if d.get(k) is not None), default=None) for k in keys}
    out['flips_total'] = int(sum(d.get('flips', 0) for d in days))
    out['flips_max_per_day'] = int(max((d.get('flips', 0) for d in days), default=0))
    out['days'] = len(days)
    return out


def record(name, rec):
    """Merge one configuration's record into gpurun_out/parity_r02.json."""
    os.makedirs(os.path.dirname(OUT_JSON), exist_ok=True)
    try:
        with open(OUT_JSON) as fobj:
            allrec = json.load(fobj)
    except Exception:
        allrec = {}
    allrec[name] = rec
    with open(OUT_JSON, 'w') as fobj:
        json.dump(allrec, fobj, indent=1, sort_keys=True)


# ---- configurations -----------------------------------------------------------------
def site_wind(gpu, tmpdir, site):
    wind, days = gpu.PM.get_wind_data(H.write_wind_file(tmpdir, site), 30, H.SITES[site])
    return wind, days


def run_probability(gpu, wind_data, days, model, pool, name, keep_per_day=True):
    """Probability model, every day: Run.solve (default options) against cs_oracle.get_solutions fed
    ORACLE-built kernels, streamed day by day (CalcSol.py:189-201)."""
    from oracle import cs_oracle as CO
    rad_res = model[-1]
    D = 2 * rad_res + 1
    t0 = time.time()
    pmfs = oracle_kernels(wind_data, days, model, None, pool)
    t_k = time.time() - t0
    w = gpu.Run.stack_wind(wind_data, days)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        res = gpu.Run.solve(w, len(days), *model, prob_model=True, want_coo=False, want_dense=True, keep_pre=True)
    recs = []
    try:
        radii_ref = [p.shape[0] // 2 for p in pmfs]
        ms = [max(p.shape[0] for p in pmfs)] * 2
        first = H.recentre(pmfs[0], rad_res)
        recs.append(day_record(res.pre(0), first.toarray(), res.dense(0), first.toarray()))
        hat = CO.fft2(first, ms)
        flags_ref = []
        t0 = time.time()
        for n in range(1, len(days)):
            CO.fftconv2(hat, pmfs[n].tocsr())
            A, flag = CO.ifft2_dense(hat, [D, D])
            dom = A[:D, :D]
            thr = CO.r_small_vals(sparse.coo_matrix(dom), prob_model=True).toarray()
            recs.append(day_record(res.pre(n), dom, res.dense(n), thr))
            flags_ref.append(bool(flag))
            if flag:
                buf = np.zeros(hat.shape)
                buf[:D, :D] = dom
                hat = CO.sfft.fft2(buf, workers=CO.WORKERS)
        t_c = time.time() - t0
        out = {'summary': summarise(recs), 'flags_equal': res.flags()[1:] == flags_ref, 'flags': [int(f) for f in flags_ref],
               'radii_equal': res.radii() == radii_ref, 'radii': radii_ref, 'P': res.P, 'P_ref': D + ms[0] // 2, 'fft_len': res.N,
               'window_steps': res.window_steps(), 'oracle_kernel_s': round(t_k, 2), 'oracle_chain_s': round(t_c, 2)}
        if keep_per_day:
            out['per_day'] = recs
    finally:
        res.close()
    record(name, out)
    return out


def run_population(gpu, wind_data, days, model, r_dur, r_number, r_start, pool, name):
    """Population model, every day (CalcSol.py:205-324), pre-threshold cohort sums and thresholded outputs."""
    from oracle import cs_oracle as CO
    rad_res = model[-1]
    D = 2 * rad_res + 1
    pmfs = oracle_kernels(wind_data, days, model, r_start, pool)
    ms = [max(p.shape[0] for p in pmfs)] * 2
    r_spread = [H.recentre(p, rad_res).tocsr() for p in pmfs[:r_dur]]
    det = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pop = CO.get_populations(r_spread, pmfs, days, len(days), D, ms, r_dur, r_number, lambda day: 1.0 / r_dur, details=det)
        w = gpu.Run.stack_wind(wind_data, days)
        res = gpu.Run.solve(w, len(days), *model, prob_model=False, r_dur=r_dur, r_number=r_number, r_dist=[1.0 / r_dur] * r_dur,
                            r_start=r_start, want_coo=False, want_dense=True, keep_pre=True)
    recs = []
    try:
        for n in range(len(days)):
            rec = day_record(res.pre(n), det['pre'][n], res.dense(n), pop[n].toarray(), prob=False)
            # the same difference in probability units (per released wasp)
            rec['max_abs_per_wasp'] = rec['max_abs'] / r_number
            recs.append(rec)
        main_flags = res.flags()[r_dur:]
        out = {'summary': summarise(recs), 'main_flags_equal': main_flags == [bool(f) for f in det['flags']],
               'flags': [int(f) for f in det['flags']], 'radii_equal': res.radii() == [p.shape[0] // 2 for p in pmfs],
               'r_number': r_number, 'r_dur': r_dur, 'P': res.P, 'fft_len': res.N, 'per_day': recs}
        out['summary']['max_abs_per_wasp'] = max(r['max_abs_per_wasp'] for r in recs)
        out['cohort_flags'] = [res.cohort_flags(n, min(n, r_dur - 1)) for n in range(len(days))] if r_dur > 1 else None
    finally:
        res.close()
    record(name, out)
    return out


def c5_cells(rad_res, n_side=32, stride=6):
    """Sample cells: a lattice around the release cell (where the population lives) plus the four corners."""
    stride = max(1, min(stride, rad_res // (n_side // 2)))
    o = rad_res - (n_side // 2) * stride
    g = [(o + i * stride, o + j * stride) for i in range(n_side) for j in range(n_side)]
    D = 2 * rad_res + 1
    g[:4] = [(0, 0), (0, D - 1), (D - 1, 0), (D - 1, D - 1)]
    return np.array(g, dtype=np.int32)


def run_c5(gpu, wind_data, days, proposals, ids, pool, name, rad_dist=10000.0, rad_res=400, r_number=130000.0):
    """Likelihood batch at the real Kalbar size: proposals through batch.solve_batch (pkb_solve_batch) against
    one oracle population solve per proposal, at the sample cells."""
    from parasitoids_b200 import batch
    cells = c5_cells(rad_res)
    w = gpu.Run.stack_wind(wind_data, days)
    props = np.asarray(proposals)[ids]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        got = batch.solve_batch(w, props, cells, len(days), rad_dist, rad_res, prob_model=False, r_dur=1, r_number=r_number)
    tasks = [(props[i], wind_data, list(days), rad_dist, rad_res, 1, r_number, None, cells) for i in range(len(ids))]
    ref = pool.map(oracle_proposal, tasks, chunksize=1)
    recs = []
    for i, (thr, pre, shapes, flags) in enumerate(ref):
        only = (got[i] != 0) != (thr != 0)
        both = ~only
        d = np.abs(np.where(both, got[i] - thr, 0.0))
        recs.append({'proposal': int(ids[i]), 'max_abs': float(d.max()), 'max_abs_per_wasp': float(d.max() / r_number),
                     'rel_l1': float(d.sum() / max(np.abs(thr).sum(), 1e-300)), 'flips': int(only.sum()),
                     'flip_vals': [float(v) for v in np.where(got[i] != 0, got[i], thr)[only][:8]],
                     'ref_max': float(np.abs(thr).max()), 'kernel_side_min_max': [min(shapes), max(shapes)],
                     'flagged_days': int(sum(flags)), 'n_periods': int(round(props[i][13])), 'corr': float(props[i][8]),
                     'mu_r': float(props[i][14])})
    out = {'proposals': recs, 'cells': int(len(cells)), 'r_number': r_number,
           'summary': {'max_abs': max(r['max_abs'] for r in recs), 'max_abs_per_wasp': max(r['max_abs_per_wasp'] for r in recs),
                       'rel_l1': max(r['rel_l1'] for r in recs), 'flips_total': sum(r['flips'] for r in recs)}}
    record(name, out)
    return out
