"""The Bayes drivers' side of the hot path: the likelihood projection (Bayes_funcs.py:20-180)
and the local day-0 spread kernel with its extra chain day (Bayes_Run.py:245-296, Bayes_MAP.py:247-277, N4).
Golden arrays come from the reference's own code (oracle/make_golden.py: make_bayes_funcs, make_sprd)."""
import os
import sys
import warnings

import numpy as np
import pytest
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import helpers as H        # noqa: E402
from oracle import bf_oracle as BO, cs_oracle as CO, pm_oracle as PO        # noqa: E402
from oracle.make_golden import bayes_locinfo        # noqa: E402  (synthetic LocInfo, pure python)


def _models(g, tag):
    return [sparse.csr_matrix(m) for m in g[tag + '_model']]


# ---- oracle against the reference's outputs (CPU) ------------------------------------
@pytest.mark.parametrize('tag,D,nd', [('a', 81, 8), ('b', 61, 30)])
def test_oracle_projection_matches_reference(tag, D, nd):
    g = H.load('bayes_funcs')
    loc = bayes_locinfo(np.random.default_rng(7), D, nd)
    rel, sen = BO.popdensity_to_emergence(_models(g, tag), loc)
    for i in range(len(rel)):
        assert np.array_equal(rel[i], g['%s_rel%d' % (tag, i)])
        assert np.allclose(sen[i], g['%s_sen%d' % (tag, i)], rtol=1e-15, atol=0)
    assert np.array_equal(BO.popdensity_grid(_models(g, tag), loc), g[tag + '_grid'])


def test_oracle_sprd_kernel_matches_reference():
    g = H.load('sprd')
    for i in range(int(g['ncases'])):
        a = g['c%d_args' % i]
        got = PO.sprd_kernel(a[0] / a[1], a[2:5], a[5:8], a[8])
        ref = g['c%d_sprd' % i]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-15
        assert abs(got.sum() - 1) < 1e-14 or got.sum() > 1


# ---- device ------------------------------------------------------------------------------
@pytest.mark.parametrize('tag,D,nd', [('a', 81, 8), ('b', 61, 30)])
def test_projection_matches_reference(pkb, tag, D, nd):
    """popdensity_to_emergence / popdensity_grid through the device projection against the arrays the
    reference's Bayes_funcs returned for the same model and LocInfo: equal to the last bit (the device sums in
    numpy's order), for the drop-in functions and for a batch of scaled copies of the model."""
    from parasitoids_b200 import Bayes_funcs as BF
    g = H.load('bayes_funcs')
    loc = bayes_locinfo(np.random.default_rng(7), D, nd)
    model = _models(g, tag)
    rel, sen = BF.popdensity_to_emergence(model, loc)
    grid = BF.popdensity_grid(model, loc)
    for i in range(len(rel)):
        assert rel[i].shape == g['%s_rel%d' % (tag, i)].shape
        assert np.array_equal(rel[i], g['%s_rel%d' % (tag, i)])
        assert np.array_equal(sen[i], g['%s_sen%d' % (tag, i)])
    assert np.array_equal(grid, g[tag + '_grid'])
    proj = BF.Projection(loc, nd)
    smp = proj.sample(model)
    out = proj.apply_samples(np.stack([smp, 0.5 * smp, 3.0 * smp]))
    r0, s0, g0 = proj.split(out[0])
    r2, s2, g2 = proj.split(out[2])
    assert np.array_equal(g0, g[tag + '_grid']) and np.array_equal(r0[1], g[tag + '_rel1'])
    assert np.allclose(r2[0], 3 * r0[0], rtol=1e-14) and np.allclose(s2[1], 3 * s0[1], rtol=1e-14)
    with pytest.raises(IndexError):
        BF.Projection(loc, 3)                     # the collections need more model days than that


def test_sprd_kernel_matches_reference(pkb):
    g = H.load('sprd')
    for i in range(int(g['ncases'])):
        a = g['c%d_args' % i]
        got = pkb.PM.sprd_kernel(a[0] / a[1], a[2:5], a[5:8], a[8])
        ref = g['c%d_sprd' % i]
        assert got.shape == ref.shape, (i, got.shape, ref.shape)
        assert np.abs(got - ref).max() < 1e-15, i


def _small_wind(nd=5, periods=48, seed=3):
    rng = np.random.default_rng(seed)
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        w[:, :, c] = 0.25 * np.sin(np.linspace(0, 5 + c, nd * periods)).reshape(nd, periods) + rng.normal(0, 0.05, (nd, periods))
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    return w


def _oracle_pop_with_sprd(w, nd, model, r_dur, r_number, r_start, sprd_factor):
    """Bayes_Run.py:236-296: pmf_list = [sprd] + prob_mass days, get_populations over ndays + 1, first day dropped."""
    hp, dp, dl, mu_r, n_periods, rad_dist, rad_res = model
    wind_data = {d: w[d] for d in range(nd)}
    pmfs = [sparse.coo_matrix(PO.sprd_kernel(rad_dist / rad_res, dp, dl, sprd_factor))]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pmfs += [PO.prob_mass(d, wind_data, hp, dp, dl, mu_r, n_periods, rad_dist, rad_res, r_start if d == 0 else None) for d in range(nd)]
    D = 2 * rad_res + 1
    ms = [max(p.shape[0] for p in pmfs)] * 2
    r_spread = [H.recentre(p, rad_res).tocsr() for p in pmfs[:r_dur]]
    det = {}
    pop = CO.get_populations(r_spread, pmfs, list(range(-1, nd)), nd + 1, D, ms, r_dur, r_number, lambda day: 1.0 / r_dur, details=det)
    return pop[1:], det['pre'][1:]


@pytest.mark.parametrize('r_dur', [1, 2])
def test_solve_with_leading_spread_day(pkb, r_dur):
    """Run.solve(sprd_factor=...) against the oracle's population model run the way Bayes_Run.pop_model runs it."""
    w, nd = _small_wind(), 5
    model = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, 1500.0, 30)
    r_number, r_start, f = 1000.0, 0.3, 0.25
    ref, ref_pre = _oracle_pop_with_sprd(w, nd, model, r_dur, r_number, r_start, f)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        res = pkb.Run.solve(w, nd, *model, prob_model=False, r_dur=r_dur, r_number=r_number, r_dist=[1.0 / r_dur] * r_dur,
                            r_start=r_start, want_coo=False, want_dense=True, keep_pre=True, sprd_factor=f)
    try:
        assert res.ndays == nd
        for d in range(nd):
            assert np.abs(res.pre(d) - ref_pre[d]).max() <= 1e-10, 'pre-threshold day %d' % d
            H.assert_thresholded_parity(res.dense(d) / r_number, ref[d].toarray() / r_number, what='day %d' % d, max_abs=1e-13)
    finally:
        res.close()


def test_batch_with_spread_day_and_projection(pkb):
    """batch.solve_batch with a per-proposal sprd_factor and the device projection: against Run.solve per proposal
    (sampled at the same cells) and against the oracle's projection of the oracle's populations."""
    from parasitoids_b200 import batch, Bayes_funcs as BF
    w, nd = _small_wind(nd=8), 8
    rad_dist, rad_res = 1500.0, 30
    D = 2 * rad_res + 1
    loc = bayes_locinfo(np.random.default_rng(11), D, nd)
    proj = BF.Projection(loc, nd)
    base = np.array([1.263, 3.913, 7.302, 2.614, 23.999, 2.350, 171.82, 144.58, 0.253, 7.096, 7.260, 0.0, 1.0, 2, 1.179])
    props = np.tile(base, (3, 1))
    props[:, 6] *= [1.0, 0.8, 1.2]
    props[:, 8] = [0.253, -0.3, 0.0]
    sf = np.array([0.1, 0.6, 0.0])
    kw = dict(prob_model=False, r_dur=1, r_number=1000.0, r_start=0.3)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        rows = batch.solve_batch(w, props, None, nd, rad_dist, rad_res, sprd_factor=sf, projection=proj, **kw)
        smp = batch.solve_batch(w, props, proj.cells, nd, rad_dist, rad_res, sprd_factor=sf, **kw)
    assert rows.shape == (3, proj.nrows) and smp.shape == (3, nd, proj.cells.shape[0])
    assert np.array_equal(rows, proj.apply_samples(smp))
    for b in range(3):
        hp, dp, dl, mu_r, n_periods = batch.unpack_proposal(props[b])
        ref, _ = _oracle_pop_with_sprd(w, nd, (hp, dp, dl, mu_r, n_periods, rad_dist, rad_res), 1, 1000.0, 0.3, sf[b])
        rel, sen = BO.popdensity_to_emergence(ref, loc)
        grid = BO.popdensity_grid(ref, loc)
        grel, gsen, ggrid = proj.split(rows[b])
        for got, want in list(zip(grel, rel)) + list(zip(gsen, sen)) + [(ggrid, grid)]:
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max())
