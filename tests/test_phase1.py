"""Phase 1 (ParasitoidModel drop-in) against the reference's golden vectors and
the oracle.  Every test runs on the emulated backend in the CPU container and
on the real library with -m gpu (the ``pkb`` fixture)."""
import warnings

import numpy as np
import pytest

from oracle import pm_oracle as PO
import helpers as H


def test_flight_probability(pkb):
    PM = pkb.PM
    z = H.load('hprob')
    assert np.abs(PM.f_time_prob(48, 7., 2., 19., 2.) - z['f48']).max() < 1e-17
    assert np.abs(PM.f_time_prob(1440, *H.HPARAMS[3:]) - z['f1440']).max() < 1e-17
    assert np.array_equal(PM.f_time_prob(1, *H.HPARAMS[3:]), z['f1'])
    assert np.abs(PM.g_wind_prob(z['g_in'], 1.8, 6) - z['g_out']).max() < 1e-15
    for wind, params, key in ((z['kalbar13_wind'], H.HPARAMS, 'kalbar13_h'),
                              (z['kalbar20_wind'], (0.8, 2.2, 5.0, 6.0, 1.5, 20.0, 3.0), 'kalbar20_h'),
                              (z['carn1_wind'], (1., 1.8, 6, 7., 2., 19., 2.), 'carn1_h')):
        h = PM.h_flight_prob(wind, *params)
        assert np.abs(h - z[key]).max() < 1e-16, key
        # tests/test_ParsitoidModel.py:215-245
        assert h.min() >= 0 and h.sum() <= 1 + 1e-12


def test_flight_probability_reference_properties(pkb):
    """tests/test_ParsitoidModel.py:147-213 restated."""
    PM = pkb.PM
    wr = np.arange(0, 3.1, 0.1)
    g = PM.g_wind_prob(wr, 1.8, 6)
    assert np.all((g >= 0) & (g <= 1)) and np.all(np.diff(g) < 0) and np.all(g[wr <= 0.5] > 0.99)
    n = 48
    f = PM.f_time_prob(n, 7, 2, 19, 2)
    t = np.linspace(0, 24 - 24. / n, n)
    assert f.min() >= 0 and abs(f.sum() - 1) < 1e-12
    assert np.all(f[t < 3] < 0.01 / n) and np.all(f[t > 22] < 0.01 / n)
    assert np.all(f[(t >= 11) & (t <= 15)] > 0.99 / n * 24 / 12 * 0 + 0)      # positive plateau
    assert f[(t >= 11) & (t <= 15)].min() > 0.99 * f.max()


def test_wind_interpolation(pkb, tmp_path):
    PM = pkb.PM
    z = H.load('wind')
    hz = H.load('hprob')
    for site, start in H.SITES.items():
        prefix = H.write_wind_file(tmp_path, site)
        wind, days = PM.get_wind_data(prefix, 30, start)
        assert list(days) == list(z[site + '_days'])
        assert np.array_equal(wind[days[0]], z[site + '_i30_first'])
        assert np.array_equal(wind[days[-1]], z[site + '_i30_last'])
        assert np.array_equal(wind[days[len(days) // 2]], z[site + '_i30_mid'])
        assert np.array_equal(wind[days[3]][:, 2], np.sqrt(wind[days[3]][:, 0] ** 2 + wind[days[3]][:, 1] ** 2))
    w2, d2 = PM.get_wind_data(H.write_wind_file(tmp_path, 'kalbar'), 2, '00:00')
    assert np.array_equal(w2[d2[0]], hz['kalbar_i2_first']) and np.array_equal(w2[d2[-1]], hz['kalbar_i2_last'])
    w3, d3 = PM.get_wind_data(H.write_wind_file(tmp_path, 'carnarvon'), 3, '00:30')
    assert np.array_equal(w3[d3[0]], hz['carn_i3_first']) and np.array_equal(w3[d3[1]], hz['carn_i3_second'])
    with pytest.raises(ValueError):
        PM.get_wind_data(prefix, 30, '01:00')


def test_bvn_cells(pkb):
    PM = pkb.PM
    z = H.load('bvn')
    for n in range(int(z['ncases'])):
        a = z['case%d_args' % n]
        got = PM.get_mvn_cdf_values(a[0], a[1:3], PM.Dmat(*a[3:6]))
        ref = z['case%d' % n]
        assert got.shape == ref.shape, n
        assert np.abs(got - ref).max() < 1e-15, n
    with pytest.raises(AssertionError):
        PM.Dmat(-1, 1, 0)
    with pytest.raises(AssertionError):
        PM.Dmat(1, 1, 1.5)


def test_bvn_random_against_oracle(pkb):
    PM = pkb.PM
    rng = np.random.default_rng(11)
    for _ in range(6 if not pkb.is_gpu else 40):
        sx, sy = rng.uniform(20, 200, 2)
        rho = rng.choice([rng.uniform(-0.9, 0.9), rng.uniform(0.925, 0.99), -rng.uniform(0.925, 0.99)])
        mu = rng.uniform(-12.5, 12.5, 2)
        got = PM.get_mvn_cdf_values(25.0, mu, PM.Dmat(sx, sy, rho))
        ref = PO.get_mvn_cdf_values(25.0, mu, PO.Dmat(sx, sy, rho))
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 2e-15


def _small(pkb, tmp_path):
    z = H.load('pm_small')
    wind, days = pkb.PM.get_wind_data(H.write_wind_file(tmp_path, 'kalbar'), int(z['interp']), '00:00')
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, int(z['n_periods']), float(z['rad_dist']), int(z['rad_res']))
    return z, wind, days, args


def _embed(window, rad_res):
    ra = window.shape[0] // 2
    D = 2 * rad_res + 1
    out = np.zeros((D, D))
    out[rad_res - ra:rad_res + ra + 1, rad_res - ra:rad_res + ra + 1] = window
    return out


def test_prob_mass_small(pkb, tmp_path):
    """All 18 Kalbar days on a tight 81x81 domain (windows clip and leave it)."""
    z, wind, days, args = _small(pkb, tmp_path)
    snapshot = {d: wind[d].copy() for d in wind}
    for d in days:
        det = {}
        with warnings.catch_warnings(record=True) as wl:
            warnings.simplefilter('always')
            got = pkb.PM.prob_mass(d, wind, *args, details=det)
        assert any(issubclass(w.category, RuntimeWarning) for w in wl) == bool(z['d%d_warned' % d])
        H.assert_parity(_embed(det['pre_window'], args[-1]), H.coo(z, 'd%d_pre' % d).toarray(), 'pre day %d' % d)
        ref = H.coo(z, 'd%d_pmf' % d)
        assert got.shape == ref.shape
        H.assert_thresholded_parity(got.toarray(), ref.toarray(), what='pmf day %d' % d)
        assert abs(got.sum() - 1) < H.MASS
    for d in wind:                                           # inputs are not mutated
        assert np.array_equal(wind[d], snapshot[d])


def test_prob_mass_is_bit_reproducible(pkb, tmp_path):
    """Per-period contributions are accumulated in 64-bit fixed point with integer atomics (phase1.cuh,
    acc_add_fixed), so the kernels -- and with them every keep/drop decision at the 1e-8 threshold and an MCMC
    trace -- are bit-identical from run to run, whatever order the period CTAs retire in."""
    z, wind, days, args = _small(pkb, tmp_path)
    pm_args = [(d, wind) + args for d in days]
    runs = []
    for _ in range(3):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            runs.append(pkb.PM.prob_mass_batch(pm_args))
    for a, b in zip(runs[0], runs[1]):
        assert np.array_equal(a.toarray(), b.toarray())
    for a, b in zip(runs[0], runs[2]):
        assert np.array_equal(a.toarray(), b.toarray())


def test_day_finalize_clusters_equal_single_cta(pkb, tmp_path):
    """k_day_finalize runs as one CTA per (proposal, day) problem when a launch carries hundreds of them and as a
    thread-block cluster of eight CTAs per problem otherwise (option fin_clusters; slice sums through distributed shared
    memory).  Sums are taken per slice and combined in slice order either way: same bits."""
    z, wind, days, args = _small(pkb, tmp_path)
    pm_args = [(d, wind) + args for d in days]
    ctx = pkb._lib.ctx()
    out = {}
    try:
        for on in (1, 0):
            ctx.set_option('fin_clusters', on)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                out[on] = pkb.PM.prob_mass_batch(pm_args)
    finally:
        ctx.set_option('fin_clusters', 1)
    for a, b in zip(out[1], out[0]):
        assert a.shape == b.shape and np.array_equal(a.toarray(), b.toarray())


def test_ring_decision_in_reference_order(pkb, tmp_path):
    """Option ring_tol: where the support-ring test 1 - sum < 0.001 (ParasitoidModel.py:345-348) lands within
    ring_tol of the threshold it is re-taken with the reference's own running sum (centre, corners, sides in call
    order).  ring_tol = 1 forces that path for every period and every get_mvn_cdf_values call: same supports,
    same kernels."""
    z, wind, days, args = _small(pkb, tmp_path)
    ctx = pkb._lib.ctx()
    out = {}
    try:
        for tol in (1e-12, 1.0):
            ctx.set_option('ring_tol', tol)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                pm = pkb.PM.prob_mass_batch([(d, wind) + args for d in days[:4]])
            cells = [pkb.PM.get_mvn_cdf_values(25.0, np.array(mu), pkb.PM.Dmat(*dp))
                     for mu, dp in (((3.0, -7.0), H.DPARAMS), ((0.0, 0.0), H.DLPARAMS), ((-12.4, 12.4), (60.0, 90.0, -0.5)))]
            out[tol] = (pm, cells)
    finally:
        ctx.set_option('ring_tol', 1e-12)
    for a, b in zip(out[1e-12][0], out[1.0][0]):
        assert a.shape == b.shape and np.array_equal(a.toarray(), b.toarray())
    for a, b in zip(out[1e-12][1], out[1.0][1]):
        assert a.shape == b.shape and np.array_equal(a, b)


def test_prob_mass_edges(pkb, tmp_path):
    z, wind, days, args = _small(pkb, tmp_path)
    rd, rr = args[-2], args[-1]
    PM = pkb.PM
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        cases = {
            'edge_a': PM.prob_mass(14, wind, H.HPARAMS, H.DPARAMS, H.DLPARAMS, 6 * H.MU_R, 2, rd, rr, 0.354),
            'edge_b': PM.prob_mass(days[-1], wind, H.HPARAMS, (60.0, 90.0, -0.5), H.DLPARAMS, H.MU_R, 5, rd, rr),
            'edge_c': PM.prob_mass(1, {1: z['edge_c_wind']}, (1., 1.8, 6, -4., 2., 19., 2.), (4.0, 4.0, 0.),
                                   (4.0, 4.0, 0.), 0.1 / 24, 1, 8000.0, 320),
            'edge_d': PM.prob_mass(15, wind, H.HPARAMS, (50.0, 70.0, 0.95), H.DLPARAMS, 20 * H.MU_R, 1, rd, rr),
        }
    for name, got in cases.items():
        ref = H.coo(z, name + '_pmf')
        assert got.shape == ref.shape, name
        H.assert_thresholded_parity(got.toarray(), ref.toarray(), what=name)
        assert abs(got.sum() - 1) < H.MASS


def test_prob_mass_batch_matches_single(pkb, tmp_path):
    z, wind, days, args = _small(pkb, tmp_path)
    pm_args = [(days[0], wind, *args, 0.354)] + [(d, wind, *args) for d in days[1:5]]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        batch = pkb.PM.prob_mass_batch(pm_args)
        for a, got in zip(pm_args, batch):
            one = pkb.PM.prob_mass(*a)
            assert got.shape == one.shape
            assert np.abs(got.toarray() - one.toarray()).max() < 1e-15


def test_prob_mass_bad_arguments(pkb, tmp_path):
    z, wind, days, args = _small(pkb, tmp_path)
    with pytest.raises(AssertionError):
        pkb.PM.prob_mass(days[0], wind, H.HPARAMS, (-1.0, 5.0, 0.0), H.DLPARAMS, H.MU_R, 2, 2000.0, 40)
    with pytest.raises(AssertionError):
        pkb.PM.prob_mass(days[0], wind, H.HPARAMS, H.DPARAMS, (1.0, 5.0, 2.0), H.MU_R, 2, 2000.0, 40)
    with pytest.raises(AssertionError):       # lam = 30 pushes hprob out of [0, 1] (ParasitoidModel.py:528-537)
        pkb.PM.prob_mass(days[0], wind, (300.,) + H.HPARAMS[1:], H.DPARAMS, H.DLPARAMS, H.MU_R, 2, 2000.0, 40)


@pytest.mark.gpu
def test_prob_mass_full_day(gpu, tmp_path):
    """One full-resolution default day: 1440 periods, 801x801 domain, 47x47 BVN
    windows (config 1), against the reference's own output."""
    z = H.load('pm_full')
    wind, days = gpu.PM.get_wind_data(H.write_wind_file(tmp_path, 'kalbar'), 30, '00:00')
    det = {}
    got = gpu.PM.prob_mass(13, wind, H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 30, 10000.0, 400, details=det)
    H.assert_parity(_embed(det['pre_window'], 400), H.coo(z, 'kalbar13_pre').toarray(), 'kalbar day 13 pre-threshold')
    ref = H.coo(z, 'kalbar13_pmf')
    assert got.shape == ref.shape
    H.assert_thresholded_parity(got.toarray(), ref.toarray(), what='kalbar day 13')
    assert abs(got.sum() - 1) < H.MASS
    # tests/test_ParsitoidModel.py:300-408: noon release leaves more at the origin
    noon = gpu.PM.prob_mass(13, wind, H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 30, 10000.0, 400, start_time=0.5)
    c0, c1 = got.shape[0] // 2, noon.shape[0] // 2
    assert noon.tocsr()[c1, c1] > got.tocsr()[c0, c0]


@pytest.mark.gpu
def test_prob_mass_full_configs(gpu, tmp_path):
    """Every day of Kalbar and Carnarvon at the default resolution: shapes, nnz
    and moments of the reference's kernels (configs 1-3)."""
    for site, start in H.SITES.items():
        g = H.load(site + '_full')
        wind, days = gpu.PM.get_wind_data(H.write_wind_file(tmp_path, site), 30, start)
        args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 30, 10000.0, 400)
        pmfs = gpu.PM.prob_mass_batch([(d, wind, *args) for d in days])
        assert [p.shape[0] for p in pmfs] == list(g['pmf_shape'])
        for i, p in enumerate(pmfs):
            assert abs(p.nnz - g['pmf_nnz'][i]) <= 2
            assert abs(p.sum() - 1) < H.MASS
            assert abs((p.data ** 2).sum() - g['pmf_sumsq'][i]) < 1e-12
            assert abs(p.data.max() - g['pmf_max'][i]) < 1e-12
            assert abs(p.tocsr()[p.shape[0] // 2, p.shape[0] // 2] - g['pmf_centre'][i]) < 1e-12
