"""The CPU oracle (oracle/) against the golden vectors produced by the
reference's own code (oracle/make_golden.py) and against the reference's
known-answer tests.  CPU only."""
import warnings

import numpy as np
import pytest
from scipy import signal, sparse

from oracle import pm_oracle as PO
from oracle import cs_oracle as CO
import helpers as H


def test_bvu_against_scipy_genz():
    """Genz BVU restatement vs SciPy's compiled BVU."""
    from scipy.special._ufuncs import _bivariate_normal_cdf as scipy_bvu
    rng = np.random.default_rng(5)
    h = rng.normal(0, 1.5, 3000)
    k = rng.normal(0, 1.5, 3000)
    for r in (0.0, -0.2, 0.253, 0.5, -0.74, 0.925, -0.925, 0.93, 0.95, -0.95, 0.99, -0.99, 0.999, -0.9999, 1.0, -1.0):
        ref = scipy_bvu(h, k, r)
        assert np.abs(PO.bvu(h, k, r) - ref).max() < 5e-15, r


def test_bvn_cells_golden():
    z = H.load('bvn')
    for n in range(int(z['ncases'])):
        a = z['case%d_args' % n]
        got = PO.get_mvn_cdf_values(a[0], a[1:3], PO.Dmat(*a[3:6]))
        ref = z['case%d' % n]
        assert got.shape == ref.shape, n
        assert np.abs(got - ref).max() < 5e-16, n


def test_bvn_reference_properties():
    """tests/test_ParsitoidModel.py:247-296 restated."""
    S1, S2 = PO.Dmat(4, 4, 0.5), PO.Dmat(10, 10, -0.5)
    c1 = PO.get_mvn_cdf_values(2, np.array([0., 0.]), S1)
    c2 = PO.get_mvn_cdf_values(2, np.array([0., 0.]), S2)
    assert 0.99 < c1.sum() < 1 and 0.99 < c2.sum() < 1
    assert c2.size > c1.size
    m1, m2 = c1.shape[0] // 2, c2.shape[0] // 2
    assert c1[:m1, :m1].sum() < c1[:m1, m1 + 1:].sum()        # rho > 0: mass along y = x
    assert c2[:m2, :m2].sum() > c2[:m2, m2 + 1:].sum()
    assert np.unravel_index(c1.argmax(), c1.shape) == (m1, m1)


def test_hprob_golden():
    z = H.load('hprob')
    assert np.array_equal(PO.f_time_prob(48, 7., 2., 19., 2.), z['f48'])
    assert np.abs(PO.f_time_prob(1440, *H.HPARAMS[3:]) - z['f1440']).max() < 1e-18
    assert np.array_equal(PO.g_wind_prob(z['g_in'], 1.8, 6), z['g_out'])
    assert np.abs(PO.h_flight_prob(z['kalbar13_wind'], *H.HPARAMS) - z['kalbar13_h']).max() < 1e-18
    assert np.abs(PO.h_flight_prob(z['kalbar20_wind'], 0.8, 2.2, 5.0, 6.0, 1.5, 20.0, 3.0) - z['kalbar20_h']).max() < 1e-18
    assert np.abs(PO.h_flight_prob(z['carn1_wind'], 1., 1.8, 6, 7., 2., 19., 2.) - z['carn1_h']).max() < 1e-18


def test_wind_interpolation_golden(tmp_path):
    """get_wind_data restated (ParasitoidModel.py:136-227) against the reference's interpolated series."""
    z = H.load('wind')
    for site, start in H.SITES.items():
        wind, days = PO.get_wind_data(H.write_wind_file(tmp_path, site), 30, start)
        assert list(days) == list(z[site + '_days'])
        assert np.array_equal(wind[days[0]], z[site + '_i30_first'])
        assert np.array_equal(wind[days[-1]], z[site + '_i30_last'])
        assert np.array_equal(wind[days[len(days) // 2]], z[site + '_i30_mid'])


def _small_wind(tmp_path):
    z = H.load('pm_small')
    wind, days = PO.get_wind_data(H.write_wind_file(tmp_path, 'kalbar'), int(z['interp']), '00:00')
    return z, wind, days


def test_prob_mass_small_golden(tmp_path):
    z, wind, days = _small_wind(tmp_path)
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, int(z['n_periods']), float(z['rad_dist']), int(z['rad_res']))
    for d in days[:6]:
        det = {}
        with warnings.catch_warnings(record=True) as wl:
            warnings.simplefilter('always')
            got = PO.prob_mass(d, wind, *args, details=det)
        assert bool(wl) == bool(z['d%d_warned' % d])
        assert np.abs(det['pmf_pre'] - H.coo(z, 'd%d_pre' % d).toarray()).max() < 1e-15
        ref = H.coo(z, 'd%d_pmf' % d)
        assert got.shape == ref.shape
        assert np.abs(got.toarray() - ref.toarray()).max() < 1e-15
        assert abs(got.sum() - 1) < H.MASS


def test_prob_mass_edges_golden(tmp_path):
    z, wind, days = _small_wind(tmp_path)
    rd, rr = float(z['rad_dist']), int(z['rad_res'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        a = PO.prob_mass(14, wind, H.HPARAMS, H.DPARAMS, H.DLPARAMS, 6 * H.MU_R, 2, rd, rr, 0.354)
        b = PO.prob_mass(days[-1], wind, H.HPARAMS, (60.0, 90.0, -0.5), H.DLPARAMS, H.MU_R, 5, rd, rr)
        c = PO.prob_mass(1, {1: z['edge_c_wind']}, (1., 1.8, 6, -4., 2., 19., 2.), (4.0, 4.0, 0.), (4.0, 4.0, 0.),
                         0.1 / 24, 1, 8000.0, 320)
        d = PO.prob_mass(15, wind, H.HPARAMS, (50.0, 70.0, 0.95), H.DLPARAMS, 20 * H.MU_R, 1, rd, rr)
    for got, name in ((a, 'edge_a'), (b, 'edge_b'), (c, 'edge_c'), (d, 'edge_d')):
        ref = H.coo(z, name + '_pmf')
        assert got.shape == ref.shape, name
        assert np.abs(got.toarray() - ref.toarray()).max() < 1e-15, name


# ---- chain ------------------------------------------------------------------
def _small_pmfs():
    z = H.load('pm_small')
    days = [int(d) for d in z['days']]
    return z, days, [H.coo(z, 'd%d_pmf' % d) for d in days]


def test_chain_small_golden():
    z, days, pmfs = _small_pmfs()
    g = H.load('chain_small')
    n = 10
    days, pmfs = days[:n], pmfs[:n]
    rr = int(z['rad_res'])
    D = int(g['dom_len'])
    ms = g['max_shape']
    det = {}
    sol = [H.recentre(pmfs[0], rr)]
    CO.get_solutions(sol, pmfs, days, n, D, ms, details=det)
    assert list(det['flags']) == [bool(f) for f in g['prob_flags']]
    for i in range(1, n):
        assert np.abs(det['pre'][i - 1] - g['prob_pre'][i - 1]).max() < 1e-15
        H.assert_thresholded_parity(sol[i].toarray(), H.coo(g, 'prob%d' % i).toarray(), what='prob day %d' % i,
                                    max_abs=1e-14, l1=1e-12)
    det = {}
    pop = CO.get_populations([H.recentre(pmfs[0], rr).tocsr()], pmfs, days, n, D, ms, 1, 130000, lambda d: 1.0, details=det)
    assert list(det['flags']) == [bool(f) for f in g['pop1_flags']]
    for i in range(n):
        H.assert_thresholded_parity(pop[i].toarray(), H.coo(g, 'pop1_%d' % i).toarray(), what='pop1 day %d' % i,
                                    max_abs=1e-14, l1=1e-12)
    r_dur = 3
    r_spread = [H.recentre(pmfs[i], rr).tocsr() for i in range(r_dur)]
    pop = CO.get_populations(r_spread, pmfs, days, n, D, ms, r_dur, 40000, lambda d: 1. / r_dur)
    for i in range(n):
        H.assert_thresholded_parity(pop[i].toarray(), H.coo(g, 'pop3_%d' % i).toarray(), what='pop3 day %d' % i,
                                    max_abs=1e-14, l1=1e-12)


def test_reference_known_answers():
    """tests/test_CalcSol.py:75-139 restated for the oracle."""
    A = np.outer(range(10), range(1, 11)).astype(float)
    B = np.outer(range(4, -1, -1), range(8, -1, -2)).astype(float)
    A_hat = CO.fft2(sparse.coo_matrix(A), np.array(B.shape))
    CO.fftconv2(A_hat, sparse.csr_matrix(B))
    assert np.allclose(np.fft.ifft2(A_hat)[:10, :10].real, signal.convolve2d(A, B, 'same'))
    A_hat = CO.fft2(sparse.coo_matrix(A), np.array([16, 16]))
    CO.fftconv2(A_hat, sparse.csr_matrix(B))
    C, flag = CO.ifft2(A_hat, A.shape)
    assert np.allclose(C.toarray(), signal.fftconvolve(A, B, 'same'))


# ---------------------------------------------------------------------------
# Container-only: the restatement against the UNMODIFIED reference modules
# executed from /root/reference (absent on the GPU box -> skipped there).
def _reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference tree not present (GPU box)')
    return ref_loader


def test_oracle_vs_live_reference_phase1(tmp_path):
    """get_mvn_cdf_values, h_flight_prob and a short prob_mass run by the
    reference's own code, now, against the oracle."""
    rl = _reference()
    from oracle import pm_oracle as PO
    pm, cs, gv = rl.load()
    rng = np.random.default_rng(42)
    for _ in range(4):
        sx, sy, rho = rng.uniform(20, 200), rng.uniform(20, 200), rng.uniform(-0.7, 0.7)
        mu = rng.uniform(-12.5, 12.5, 2)
        with rl.quiet():
            ref = pm.get_mvn_cdf_values(25.0, mu, pm.Dmat(sx, sy, rho))
        got = PO.get_mvn_cdf_values(25.0, mu, PO.Dmat(sx, sy, rho))
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 5e-16
    periods = 48
    w = np.zeros((2, periods, 3))
    w[:, :, 0] = 0.4 * np.sin(np.linspace(0, 4, 2 * periods)).reshape(2, periods)
    w[:, :, 1] = 0.2 * np.cos(np.linspace(0, 3, 2 * periods)).reshape(2, periods)
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    hp = (1., 1.263, 3.913, 7.302, 2.614, 23.999, 2.350)
    with rl.quiet():
        href = pm.h_flight_prob(w[0], *hp)
    assert np.abs(PO.h_flight_prob(w[0], *hp) - href).max() < 1e-17
    wind_data = {1: w[0], 2: w[1]}
    args = (1, wind_data, hp, (171.82, 144.58, 0.253), (7.096, 7.260, 0.0), 1.179, 2, 1500.0, 30)
    import warnings
    with rl.quiet(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = pm.prob_mass(*args).toarray()
        got = PO.prob_mass(*args).toarray()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-15 and abs(got.sum() - 1) < 1e-14


def test_oracle_vs_live_reference_chain():
    """get_solutions / get_populations (with the documented back_solve fix) by
    the reference's own CalcSol, now, against the oracle."""
    rl = _reference()
    from oracle import cs_oracle as CO
    pm, cs, gv = rl.load(fixed_back_solve=True)
    rng = np.random.default_rng(9)
    D = 41
    pmfs = []
    for k in (9, 13, 11, 15, 9):
        a = rng.random((k, k))
        a[a < 0.3] = 0
        pmfs.append(sparse.coo_matrix(a / a.sum()))
    ms = [15, 15]
    days = list(range(5))

    def first():
        p = pmfs[0]
        off = D // 2 - p.shape[0] // 2
        return sparse.coo_matrix((p.data, (p.row + off, p.col + off)), shape=(D, D))
    ref, got = [first()], [first()]
    with rl.quiet():
        cs.get_solutions(ref, pmfs, days, 5, D, ms)
    CO.get_solutions(got, pmfs, days, 5, D, ms)
    for a, b in zip(ref, got):
        assert np.abs(a.toarray() - b.toarray()).max() < 1e-15
    r_spread = []
    for p in pmfs[:2]:
        off = D // 2 - p.shape[0] // 2
        r_spread.append(sparse.coo_matrix((p.data, (p.row + off, p.col + off)), shape=(D, D)).tocsr())
    import warnings
    with rl.quiet(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pref = cs.get_populations(r_spread, pmfs, days, 5, D, ms, 2, 1000.0, lambda d: 0.5)
    pgot = CO.get_populations(r_spread, pmfs, days, 5, D, ms, 2, 1000.0, lambda d: 0.5)
    for a, b in zip(pref, pgot):
        assert np.abs(a.toarray() - b.toarray()).max() < 1e-11
