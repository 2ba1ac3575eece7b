"""Parity of the BENCHMARKED paths at full size, every day, against the oracle fed
ORACLE-built kernels (VERDICT r1 "next round" item 1).  All `-m gpu`; the numbers
land in gpurun_out/parity_r02.json (committed as profiles/parity_r02.json).

  C1/C2/C3  Kalbar and Carnarvon at default resolution (801^2): every day's dense
            pre-threshold grid, flags of every ifft2 call (incl. back_solve), radii
  C4        synthetic 4097^2, all 60 days through Run.solve with DEFAULT options --
            the code path bench.py times (support-window steps, fused row passes,
            side-stream emission, spectral-resident steps)
  C5        16 of the 512 prior draws of bench.prior_proposals at the real Kalbar
            size through batch.solve_batch (own torus per proposal)

Bars (BASELINE.json north_star): max-abs 1e-10, relative L1 1e-9 on pre-threshold
grids, |sum - 1| <= 1e-12 on probability days.  Population grids are the same
cohort sums times r_number (130000 / 40000 wasps): the bar is applied to the raw
population values as well (no scaling by max|ref|); the per-wasp figure is reported
beside it.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import helpers as H        # noqa: E402
import parity_lib as PL    # noqa: E402

MODEL_801 = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 30, 10000.0, 400)


@pytest.fixture(scope='module')
def pool():
    from oracle import cs_oracle as CO
    CO.WORKERS = -1          # bit-identical results, less wall time (see cs_oracle.WORKERS)
    p = PL.make_pool()
    yield p
    p.close()
    p.join()
    CO.WORKERS = None


def _check_prob(out, what):
    s = out['summary']
    assert out['flags_equal'], what + ': boundary-flag sequence differs from the oracle'
    assert out['radii_equal'], what + ': kernel radii differ from the oracle'
    assert out['P'] == out['P_ref']
    assert s['max_abs'] <= H.MAX_ABS, '%s: pre-threshold max-abs %.3e' % (what, s['max_abs'])
    assert s['rel_l1'] <= H.REL_L1, '%s: pre-threshold relative L1 %.3e' % (what, s['rel_l1'])
    assert s['mass_err'] <= H.MASS, '%s: mass error %.3e' % (what, s['mass_err'])
    assert s['thr_max_abs'] <= H.MAX_ABS
    # keep/drop flips only for cells sitting on the 1e-8 threshold (~1e-3 per day expected)
    assert s['flips_max_per_day'] <= 3, '%s: %d support flips in one day' % (what, s['flips_max_per_day'])
    for d in out.get('per_day', []):
        if d.get('flips'):
            assert d['flip_max_rel_dist'] < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('site', ['kalbar', 'carnarvon'])
def test_full_size_probability_every_day(gpu, tmp_path, pool, site):
    """C1 / C3 probability model at 801^2: all 18 / 30 days, dense, pre-threshold."""
    wind, days = PL.site_wind(gpu, tmp_path, site)
    out = PL.run_probability(gpu, wind, days, MODEL_801, pool, 'c1_kalbar_prob' if site == 'kalbar' else 'c3_carnarvon_prob')
    _check_prob(out, site)
    g = H.load(site + '_full')       # and the flags the REFERENCE itself produced (oracle/make_golden.py)
    assert out['flags'] == [int(f) for f in g['prob_flags']]


@pytest.mark.gpu
@pytest.mark.parametrize('site', ['kalbar', 'carnarvon'])
def test_full_size_population_every_day(gpu, tmp_path, pool, site):
    """C2 / C3 population model at 801^2: every day's cohort sum before r_small_vals, the thresholded
    output, and the flag of EVERY ifft2 call (main chain and back_solve) against the reference's own
    record (`pop_flags` of the golden files, CalcSol.py:99-105,189-201 call order)."""
    wind, days = PL.site_wind(gpu, tmp_path, site)
    r_dur, r_number, r_start = (1, 130000.0, None) if site == 'kalbar' else (5, 40000.0, 0.354)
    out = PL.run_population(gpu, wind, days, MODEL_801, r_dur, r_number, r_start, pool,
                            'c2_kalbar_pop' if site == 'kalbar' else 'c3_carnarvon_pop')
    s = out['summary']
    assert out['main_flags_equal'] and out['radii_equal']
    # raw population values against the un-scaled bar; per-wasp (probability) units beside it
    assert s['max_abs'] <= H.MAX_ABS, '%s pop: max-abs %.3e wasps' % (site, s['max_abs'])
    assert s['max_abs_per_wasp'] <= 1e-14
    assert s['rel_l1'] <= H.REL_L1
    assert s['thr_max_abs'] <= H.MAX_ABS
    assert s['flips_max_per_day'] <= 3
    # flag of every ifft2 call in the reference's order: release days 1..r_dur-1 (back_solve only), then per
    # post-release day the main flag followed by the back_solve flags
    g = H.load(site + '_full')
    seq = []
    res_flags = out['flags']
    for n in range(1, len(days)):
        if n >= r_dur:
            seq.append(bool(res_flags[n - r_dur]))
        if r_dur > 1:
            seq.extend(out['cohort_flags'][n])
    assert seq == [bool(f) for f in g['pop_flags']]


@pytest.mark.gpu
def test_c4_all_days_against_oracle(gpu, pool):
    """BASELINE config 4 exactly as bench.py runs it: 60 days at 4097^2, default options."""
    import bench
    wind, wind_data, days, rad_dist, rad_res = bench.load_workload('synthetic_4097x4097_60d')
    model = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 30, rad_dist, rad_res)
    out = PL.run_probability(gpu, wind_data, days, model, pool, 'c4_synthetic_4097_60d')
    _check_prob(out, 'C4')
    assert out['P'] == 4277 and not any(out['flags'])
    assert min(out['radii']) == 66 and max(out['radii']) == 180
    assert out['window_steps'] >= 10


@pytest.mark.gpu
def test_c4_kernels_pre_threshold_against_oracle(gpu, pool):
    """Phase 1 on the C4 wind: every day's pre-threshold accumulation window against the oracle."""
    import bench
    wind, wind_data, days, rad_dist, rad_res = bench.load_workload('synthetic_4097x4097_60d')
    model = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 30, rad_dist, rad_res)
    tasks = []
    for d in days:
        sub = {k: wind_data[k] for k in (d, d + 1) if k in wind_data}
        tasks.append((d, sub, True) + model)
    ref = pool.map(PL.oracle_pre_window, tasks, chunksize=1)
    args = [gpu.PM._day_args(model[0], model[1], model[2], model[3], model[4], rad_dist, rad_res, None, i, False)
            for i in range(len(days))]
    ks = gpu.PM.build_kernels(wind, args, keep_pre=True)
    worst, recs = 0.0, []
    try:
        racc = gpu._lib.lib().pkb_kset_racc(ks.h)
        for i in range(len(days)):
            got = ks.pre(i)
            r = ref[i]
            rr = r.shape[0] // 2
            assert rr <= racc
            win = got[racc - rr:racc + rr + 1, racc - rr:racc + rr + 1]
            outside = got.copy()
            outside[racc - rr:racc + rr + 1, racc - rr:racc + rr + 1] = 0
            assert not outside.any()
            recs.append(PL.day_record(win, r))
            worst = max(worst, recs[-1]['max_abs'])
    finally:
        ks.close()
    PL.record('c4_phase1_pre_threshold', {'summary': PL.summarise(recs), 'per_day': recs})
    assert worst <= 1e-14        # sums of ~1400 cell masses of <= 4e-3 each


@pytest.mark.gpu
def test_c5_prior_draws_real_size(gpu, tmp_path, pool):
    """16 of the 512 prior draws bench.py's kalbar_batch512 workload uses, at the real Kalbar size."""
    import bench
    wind, days = PL.site_wind(gpu, tmp_path, 'kalbar')
    props = bench.prior_proposals(512)
    # a spread over the batch plus the extremes the draws contain: largest |corr|, smallest and largest mu_r,
    # smallest and largest n_periods
    ids = set(range(0, 512, 47))
    for col, fn in ((8, np.argmax), (8, np.argmin), (14, np.argmin), (14, np.argmax), (13, np.argmin), (13, np.argmax)):
        ids.add(int(fn(props[:, col])))
    ids = sorted(ids)[:16]
    out = PL.run_c5(gpu, wind, days, props, ids, pool, 'c5_prior_draws')
    s = out['summary']
    assert s['max_abs'] <= H.MAX_ABS, 'C5: max-abs %.3e wasps at the sample cells' % s['max_abs']
    assert s['rel_l1'] <= H.REL_L1
    assert s['flips_total'] <= 4
