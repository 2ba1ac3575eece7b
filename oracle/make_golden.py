"""Generate tests/golden/*.npz by running the REFERENCE's own code.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Build-container only
(needs /root/reference).  Usage:

    python -m oracle.make_golden [wind] [bvn] [hprob] [pm_small] [chain_small]
                                 [pm_full] [kalbar_full] [carnarvon_full]

With no arguments the quick sets (bvn hprob pm_small chain_small) are made.
Every array written here is an output of unmodified reference functions
(ParasitoidModel.py / CalcSol.py executed from /root/reference through
oracle/ref_loader.py, i.e. with the 3-line mvn.mvnun shim), except the
population-model cases with r_dur > 1, which use the reference plus the
documented one-line back_solve correction (ref_loader.load(True)).
"""
import os
import sys
import warnings
from multiprocessing import Pool

import numpy as np
from scipy import sparse

from . import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
DATA = os.path.join(ref_loader.REF_ROOT, 'data')

KALBAR = dict(site=os.path.join(DATA, 'kalbar'), start='00:00')
CARNARVON = dict(site=os.path.join(DATA, 'carnarvonearl'), start='00:30')

# Run.py:68-83 defaults
HPARAMS = (1., 1.263, 3.913, 7.302, 2.614, 23.999, 2.350)
DPARAMS = (171.82, 144.58, 0.253)
DLPARAMS = (7.096, 7.260, 0.000)
MU_R = 1.179

BVN_CASES = [
    # (cell_length, mu, (sig_x, sig_y, rho))
    (25.0, (3.0, -7.0), DPARAMS),
    (25.0, (0.0, 0.0), DLPARAMS),
    (2.0, (0.0, 0.0), (4.0, 4.0, 0.5)),          # tests/test_ParsitoidModel.py:258-261
    (2.0, (0.0, 0.0), (10.0, 10.0, -0.5)),       # :263-266
    (25.0, (1.0, 2.0), (120.0, 100.0, 0.6)),     # 12-point rule
    (25.0, (-12.4, 12.49), (60.0, 90.0, -0.74)),
    (25.0, (5.0, 5.0), (100.0, 80.0, 0.95)),     # |rho| >= 0.925 branch
    (25.0, (-12.4, 12.49), (60.0, 90.0, -0.93)),
    (25.0, (12.5, -12.5), (40.0, 40.0, 0.0)),    # remainder exactly on a cell edge
    (10.0, (0.3, 4.9), (35.0, 21.0, 0.29)),
]


def _save(name, **arrays):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('wrote', path, os.path.getsize(path), 'bytes')


def make_bvn():
    pm, _, _ = ref_loader.load()
    out = {}
    for n, (cell, mu, dp) in enumerate(BVN_CASES):
        S = pm.Dmat(*dp)
        out['case%d_args' % n] = np.array([cell, mu[0], mu[1], *dp])
        out['case%d' % n] = pm.get_mvn_cdf_values(cell, np.array(mu), S)
    out['ncases'] = np.array(len(BVN_CASES))
    _save('bvn', **out)


def make_hprob():
    pm, _, _ = ref_loader.load()
    out = {}
    wk, _ = pm.get_wind_data(KALBAR['site'], 30, KALBAR['start'])
    wc, _ = pm.get_wind_data(CARNARVON['site'], 30, CARNARVON['start'])
    out['kalbar13_wind'] = wk[13]
    out['kalbar13_h'] = pm.h_flight_prob(wk[13], *HPARAMS)
    out['kalbar20_h'] = pm.h_flight_prob(wk[20], 0.8, 2.2, 5.0, 6.0, 1.5, 20.0, 3.0)
    out['kalbar20_wind'] = wk[20]
    out['carn1_wind'] = wc[1]
    out['carn1_h'] = pm.h_flight_prob(wc[1], 1., 1.8, 6, 7., 2., 19., 2.)
    out['f48'] = pm.f_time_prob(48, 7., 2., 19., 2.)
    out['f1440'] = pm.f_time_prob(1440, *HPARAMS[3:])
    out['f1'] = pm.f_time_prob(1, *HPARAMS[3:])
    wr = np.arange(0, 3.1, 0.1)
    out['g_in'] = wr
    out['g_out'] = pm.g_wind_prob(wr, 1.8, 6)
    # wind interpolation fence posts (inputs of the path, ParasitoidModel.py:136-227)
    w2, d2 = pm.get_wind_data(KALBAR['site'], 2, KALBAR['start'])
    out['kalbar_i2_days'] = np.array(d2)
    out['kalbar_i2_first'] = w2[d2[0]]
    out['kalbar_i2_last'] = w2[d2[-1]]
    w3, d3 = pm.get_wind_data(CARNARVON['site'], 3, CARNARVON['start'])
    out['carn_i3_days'] = np.array(d3)
    out['carn_i3_first'] = w3[d3[0]]
    out['carn_i3_second'] = w3[d3[1]]
    _save('hprob', **out)


def make_wind():
    """Raw wind series of both sites as the reference's read_wind_file returns
    them (data/*wind.txt are inputs of the path, not code); the tests rebuild
    text files from these rows because /root/reference is absent on the GPU box."""
    pm, _, _ = ref_loader.load()
    out = {}
    for name, cfg in (('kalbar', KALBAR), ('carnarvon', CARNARVON)):
        raw, days = pm.read_wind_file(cfg['site'])
        out[name + '_days'] = np.array(days)
        out[name + '_raw'] = np.stack([raw[d] for d in days])      # (ndays, 48, 3)
        w30, _ = pm.get_wind_data(cfg['site'], 30, cfg['start'])
        # spot rows of the 1-minute interpolation, to pin get_wind_data itself
        out[name + '_i30_first'] = w30[days[0]]
        out[name + '_i30_last'] = w30[days[-1]]
        out[name + '_i30_mid'] = w30[days[len(days) // 2]]
    _save('wind', **out)


def _capture_prob_mass(pm, *args, **kw):
    """Run reference prob_mass, also capturing the dense grid it hands to
    r_small_vals (ParasitoidModel.py:605) and whether it warned."""
    captured = {}
    orig = pm.r_small_vals

    def spy(A, prob_model=False, negval=1e-8):
        captured['pre'] = sparse.coo_matrix(A).copy()
        return orig(A, prob_model=prob_model, negval=negval)

    pm.r_small_vals = spy
    try:
        with warnings.catch_warnings(record=True) as wlist:
            warnings.simplefilter('always')
            with ref_loader.quiet():
                out = pm.prob_mass(*args, **kw)
    finally:
        pm.r_small_vals = orig
    captured['warned'] = any(issubclass(w.category, RuntimeWarning) for w in wlist)
    return out, captured


def _coo_arrays(prefix, M):
    M = sparse.coo_matrix(M)
    return {prefix + '_row': M.row.astype(np.int32), prefix + '_col': M.col.astype(np.int32),
            prefix + '_val': M.data.astype(np.float64), prefix + '_shape': np.array(M.shape)}


# small configuration: 15-minute periods (interp_num=2), 30-minute flights,
# 50 m cells on a deliberately tight domain so windows clip and leave it.
SMALL = dict(interp=2, n_periods=2, rad_dist=2000.0, rad_res=40)
CHAIN_DAYS = 10


def _small_day(day):
    pm, _, _ = ref_loader.load()
    wind, days = pm.get_wind_data(KALBAR['site'], SMALL['interp'], KALBAR['start'])
    out, cap = _capture_prob_mass(pm, day, wind, HPARAMS, DPARAMS, DLPARAMS, MU_R,
                                  SMALL['n_periods'], SMALL['rad_dist'], SMALL['rad_res'])
    return day, out, cap


def make_pm_small():
    pm, _, _ = ref_loader.load()
    wind, days = pm.get_wind_data(KALBAR['site'], SMALL['interp'], KALBAR['start'])
    with Pool() as pool:
        res = pool.map(_small_day, days)
    out = {'days': np.array(days), 'hparams': np.array(HPARAMS), 'dparams': np.array(DPARAMS),
           'dlparams': np.array(DLPARAMS), 'mu_r': np.array(MU_R),
           'interp': np.array(SMALL['interp']), 'n_periods': np.array(SMALL['n_periods']),
           'rad_dist': np.array(SMALL['rad_dist']), 'rad_res': np.array(SMALL['rad_res'])}
    for day, pmf, cap in res:
        out.update(_coo_arrays('d%d_pmf' % day, pmf))
        out.update(_coo_arrays('d%d_pre' % day, cap['pre']))
        out['d%d_warned' % day] = np.array(cap['warned'])
    # extra edge cases on one day -------------------------------------------------
    # (a) start_time (release at 08:30, Run.py:122) and strong drift that leaves
    #     the domain entirely for some periods (mu_r x6)
    pmf, cap = _capture_prob_mass(pm, 14, wind, HPARAMS, DPARAMS, DLPARAMS, 6 * MU_R,
                                  SMALL['n_periods'], SMALL['rad_dist'], SMALL['rad_res'], 0.354)
    out.update(_coo_arrays('edge_a_pmf', pmf))
    out.update(_coo_arrays('edge_a_pre', cap['pre']))
    out['edge_a_warned'] = np.array(cap['warned'])
    # (b) last day of the data (extrapolated flight average, :455-460), n_periods=5
    pmf, cap = _capture_prob_mass(pm, days[-1], wind, HPARAMS, (60.0, 90.0, -0.5), DLPARAMS, MU_R,
                                  5, SMALL['rad_dist'], SMALL['rad_res'])
    out.update(_coo_arrays('edge_b_pmf', pmf))
    out.update(_coo_arrays('edge_b_pre', cap['pre']))
    # (c) single-period test form (1-D wind row, tests/test_ParsitoidModel.py:315-325)
    w30, _ = pm.get_wind_data(CARNARVON['site'], 30, CARNARVON['start'])
    row = w30[1][24 * 30, :]
    out['edge_c_wind'] = row
    pmf, cap = _capture_prob_mass(pm, 1, {1: row}, (1., 1.8, 6, -4., 2., 19., 2.), (4.0, 4.0, 0.),
                                  (4.0, 4.0, 0.), 0.1 / 24, 1, 8000.0, 320)
    out.update(_coo_arrays('edge_c_pmf', pmf))
    # (d) n_periods = 1 (no flight averaging, :461-462), high correlation
    pmf, cap = _capture_prob_mass(pm, 15, wind, HPARAMS, (50.0, 70.0, 0.95), DLPARAMS, 20 * MU_R,
                                  1, SMALL['rad_dist'], SMALL['rad_res'])
    out.update(_coo_arrays('edge_d_pmf', pmf))
    out.update(_coo_arrays('edge_d_pre', cap['pre']))
    _save('pm_small', **out)


def _load_small_pmfs():
    z = np.load(os.path.join(GOLD, 'pm_small.npz'))
    days = [int(d) for d in z['days']]
    pmfs = []
    for d in days:
        shp = tuple(int(s) for s in z['d%d_pmf_shape' % d])
        pmfs.append(sparse.coo_matrix((z['d%d_pmf_val' % d],
                                       (z['d%d_pmf_row' % d], z['d%d_pmf_col' % d])), shape=shp))
    return z, days, pmfs


def _spy_chain(cs):
    """Wrap reference ifft2 to record every boundary flag and dense result."""
    rec = {'flags': [], 'pre': []}
    orig = cs.ifft2

    def spy(A_hat, Ashape):
        A, flag = orig(A_hat, Ashape)
        rec['flags'].append(bool(flag))
        rec['pre'].append(A.toarray())
        return A, flag

    cs.ifft2 = spy
    return rec


def _recentre(pmf, rad_res):
    """Run.py:454-458."""
    off = rad_res - pmf.shape[0] // 2
    D = 2 * rad_res + 1
    return sparse.coo_matrix((pmf.data, (pmf.row + off, pmf.col + off)), shape=(D, D))


def make_chain_small():
    z, days, pmfs = _load_small_pmfs()
    days, pmfs = days[:CHAIN_DAYS], pmfs[:CHAIN_DAYS]
    rad_res = int(z['rad_res'])
    D = 2 * rad_res + 1
    max_shape = np.array([max(p.shape[0] for p in pmfs)] * 2)
    out = {'max_shape': max_shape, 'dom_len': np.array(D)}

    # probability model (Run.py:450-464 + CalcSol.get_solutions)
    _, cs, _ = ref_loader.load()
    rec = _spy_chain(cs)
    modelsol = [_recentre(pmfs[0], rad_res)]
    with ref_loader.quiet():
        cs.get_solutions(modelsol, pmfs, days, len(days), D, max_shape)
    out['prob_flags'] = np.array(rec['flags'])
    for n, sol in enumerate(modelsol):
        out.update(_coo_arrays('prob%d' % n, sol))
    out['prob_pre'] = np.array(rec['pre'])

    # population model r_dur = 1 (Kalbar form, Run.py:126-138)
    _, cs, _ = ref_loader.load()
    rec = _spy_chain(cs)
    with ref_loader.quiet(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pop = cs.get_populations([_recentre(pmfs[0], rad_res).tocsr()], pmfs, days, len(days),
                                 D, max_shape, 1, 130000, lambda d: 1.0)
    out['pop1_flags'] = np.array(rec['flags'])
    for n, sol in enumerate(pop):
        out.update(_coo_arrays('pop1_%d' % n, sol))
    out['pop1_ndays'] = np.array(len(pop))

    # population model r_dur = 3, uniform emergence, reference + back_solve fix
    _, cs, _ = ref_loader.load(fixed_back_solve=True)
    rec = _spy_chain(cs)
    r_dur = 3
    r_spread = [_recentre(pmfs[i], rad_res).tocsr() for i in range(r_dur)]
    with ref_loader.quiet(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pop = cs.get_populations(r_spread, pmfs, days, len(days), D, max_shape, r_dur, 40000,
                                 lambda d: 1. / r_dur)
    out['pop3_flags'] = np.array(rec['flags'])
    for n, sol in enumerate(pop):
        out.update(_coo_arrays('pop3_%d' % n, sol))
    out['pop3_ndays'] = np.array(len(pop))
    _save('chain_small', **out)


def _full_day(args):
    which, day, start_time = args
    pm, _, _ = ref_loader.load()
    cfg = KALBAR if which == 'kalbar' else CARNARVON
    wind, days = pm.get_wind_data(cfg['site'], 30, cfg['start'])
    a = (day, wind, HPARAMS, DPARAMS, DLPARAMS, MU_R, 30, 10000.0, 400)
    if start_time is not None:
        a = a + (start_time,)
    out, cap = _capture_prob_mass(pm, *a)
    return day, out, cap


def make_pm_full():
    """One full-resolution default day with its pre-threshold grid."""
    day, pmf, cap = _full_day(('kalbar', 13, None))
    out = {}
    out.update(_coo_arrays('kalbar13_pmf', pmf))
    out.update(_coo_arrays('kalbar13_pre', cap['pre']))
    _save('pm_full', **out)


def _stats(prefix, mats):
    mats = [sparse.coo_matrix(m) for m in mats]
    arg = [int(np.argmax(m.data)) for m in mats]
    return {
        prefix + '_shape': np.array([m.shape[0] for m in mats]),
        prefix + '_nnz': np.array([m.nnz for m in mats]),
        prefix + '_sum': np.array([m.data.sum() for m in mats]),
        prefix + '_sumsq': np.array([(m.data ** 2).sum() for m in mats]),
        prefix + '_max': np.array([m.data.max() for m in mats]),
        prefix + '_argmax': np.array([[m.row[a], m.col[a]] for m, a in zip(mats, arg)]),
        prefix + '_centre': np.array([m.tocsr()[m.shape[0] // 2, m.shape[1] // 2] for m in mats]),
    }


def _make_full(which):
    pm, _, _ = ref_loader.load()
    cfg = KALBAR if which == 'kalbar' else CARNARVON
    wind, days = pm.get_wind_data(cfg['site'], 30, cfg['start'])
    r_start = None if which == 'kalbar' else 0.354
    r_dur = 1 if which == 'kalbar' else 5
    r_number = 130000 if which == 'kalbar' else 40000
    rad_res = 400
    D = 2 * rad_res + 1
    out = {'days': np.array(days)}
    # probability-model kernels (no start_time), Run.py:414-416
    with Pool() as pool:
        res = pool.map(_full_day, [(which, d, None) for d in days])
    pmfs = [r[1] for r in res]
    out.update(_stats('pmf', pmfs))
    max_shape = np.array([max(p.shape[0] for p in pmfs)] * 2)
    out['max_shape'] = max_shape
    _, cs, _ = ref_loader.load()
    rec = _spy_chain(cs)
    modelsol = [_recentre(pmfs[0], rad_res)]
    with ref_loader.quiet():
        cs.get_solutions(modelsol, pmfs, days, len(days), D, max_shape)
    out['prob_flags'] = np.array(rec['flags'])
    out.update(_stats('prob', modelsol))
    # a few full rows/columns of the last day so values, not only moments, are pinned
    last = modelsol[-1].toarray()
    out['prob_last_row400'] = last[400, :]
    out['prob_last_col380'] = last[:, 380]
    # population model, Run.py:417-421,466-481
    if r_start is not None:
        d0 = _full_day((which, days[0], r_start))
        pmfs_pop = [d0[1]] + pmfs[1:]
        out.update(_stats('pmf_pop0', [d0[1]]))
    else:
        pmfs_pop = pmfs
    max_shape_p = np.array([max(p.shape[0] for p in pmfs_pop)] * 2)
    out['max_shape_pop'] = max_shape_p
    _, cs, _ = ref_loader.load(fixed_back_solve=True)
    rec = _spy_chain(cs)
    r_spread = [_recentre(pmfs_pop[i], rad_res).tocsr() for i in range(r_dur)]
    with ref_loader.quiet(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pop = cs.get_populations(r_spread, pmfs_pop, days, len(days), D, max_shape_p, r_dur,
                                 r_number, lambda d: 1. / r_dur)
    out['pop_flags'] = np.array(rec['flags'])
    out.update(_stats('pop', pop))
    lastp = pop[-1].toarray()
    out['pop_last_row400'] = lastp[400, :]
    out['pop_last_col380'] = lastp[:, 380]
    _save(which + '_full', **out)


SPRD_CASES = [
    # (domain_info, Dparams, Dlparams, sprd_factor)
    ((10000.0, 400), DPARAMS, DLPARAMS, 0.1),            # Bayes_Run defaults (res 25 m)
    ((10000.0, 200), DPARAMS, DLPARAMS, 0.35),           # Bayes_MAP domain (res 50 m)
    ((10000.0, 400), (90.0, 60.0, -0.6), (20.0, 31.0, 0.4), 0.8),
    ((8000.0, 500), (30.0, 40.0, 0.95), (9.0, 9.0, 0.0), 0.0),
    ((2000.0, 40), DPARAMS, DLPARAMS, 1.0),
]


def _reference_sprd_source():
    """The body of the `if sprd_factor is not None:` block of Bayes_Run.pop_model (Bayes_Run.py:246-269), read from
    the reference tree at generation time and executed as it stands (Bayes_Run itself cannot be imported: pymc)."""
    import textwrap
    with open(os.path.join(ref_loader.REF_ROOT, 'Bayes_Run.py')) as fobj:
        lines = fobj.read().split('\n')
    start = next(i for i, ln in enumerate(lines) if ln.strip() == 'if sprd_factor is not None:') + 1
    end = next(i for i in range(start, len(lines)) if lines[i].strip().startswith('pmf_list = [sparse.coo_matrix(sprd)]'))
    return textwrap.dedent('\n'.join(lines[start:end]))


def make_sprd():
    pm, _, _ = ref_loader.load()
    src = _reference_sprd_source()
    out = {'ncases': np.array(len(SPRD_CASES))}
    for i, (dom, dp, dl, f) in enumerate(SPRD_CASES):
        class _P(object):
            domain_info = dom
        ns = {'np': np, 'PM': pm, 'params': _P(), 'params_ary': list(range(6)) + list(dp) + list(dl), 'sprd_factor': f}
        exec(src, ns)
        out['c%d_sprd' % i] = ns['sprd']
        out['c%d_args' % i] = np.array([dom[0], dom[1], *dp, *dl, f])
    _save('sprd', **out)


class _Days(object):
    """Stand-in for a pandas Timedelta: Bayes_funcs only reads `.days`."""
    def __init__(self, d):
        self.days = int(d)


def bayes_locinfo(rng, dom_len, ndays):
    """A synthetic LocInfo with the attributes Bayes_funcs.py:20-173 reads (Data_Import.LocInfo builds the real one
    from the field spreadsheets): two collections with emergence grids and observation dates, three sentinel
    fields, a release-field grid observed on two days."""
    import pandas as pd

    class Loc(object):
        pass
    loc = Loc()
    mid = dom_len // 2
    span = max(3, dom_len // 10)
    loc.collection_datesPR = [_Days(min(ndays - 1, 6)), _Days(ndays)]
    loc.emerg_grids = [[(int(mid + rng.integers(-span, span)), int(mid + rng.integers(-span, span))) for _ in range(n)] for n in (7, 11)]
    loc.release_DataFrames = []
    loc.sent_DataFrames = []
    for cd, obs in zip(loc.collection_datesPR, ((0, 3, 4, 9), (2, 5, 20, 24))):
        dates = [pd.Timedelta(days=cd.days + o) for o in obs for _ in range(2)]
        loc.release_DataFrames.append(pd.DataFrame({'datePR': dates}))
        loc.sent_DataFrames.append(pd.DataFrame({'datePR': dates[::2] + dates[-1:]}))
    loc.sent_ids = ['A', 'B', 'C']
    loc.field_cells = {}
    for k, fid in enumerate(loc.sent_ids):
        r0, c0 = mid + (k - 1) * span, mid - span + 2 * k
        rr, cc = np.meshgrid(np.arange(r0, r0 + 4 + k), np.arange(c0, c0 + 5))
        loc.field_cells[fid] = np.stack([rr.ravel(), cc.ravel()], axis=1)
    loc.grid_cells = np.array([(int(mid + rng.integers(-span, span)), int(mid + rng.integers(-span, span))) for _ in range(9)])
    loc.grid_obs_datesPR = [_Days(2), _Days(min(ndays, 5))]
    return loc


def make_bayes_funcs():
    """Bayes_funcs.popdensity_to_emergence / popdensity_grid (the reference module imports as it stands) on the
    small Kalbar population solution of chain_small and on a random dense model of 30 days."""
    ref_loader.load()
    bf = __import__('Bayes_funcs')
    out = {}
    rng = np.random.default_rng(42)
    for tag, D, nd in (('a', 81, 8), ('b', 61, 30)):
        model = [sparse.csr_matrix(np.where(rng.random((D, D)) < 0.3, rng.random((D, D)) * 50, 0.0)) for _ in range(nd)]
        loc = bayes_locinfo(np.random.default_rng(7), D, nd)
        rel, sen = bf.popdensity_to_emergence(model, loc)
        grid = bf.popdensity_grid(model, loc)
        out[tag + '_model'] = np.stack([m.toarray() for m in model])
        for i, (r, s_) in enumerate(zip(rel, sen)):
            out['%s_rel%d' % (tag, i)] = r
            out['%s_sen%d' % (tag, i)] = s_
        out[tag + '_grid'] = grid
    _save('bayes_funcs', **out)


def main(argv):
    if not ref_loader.available():
        sys.exit('reference tree not found at ' + ref_loader.REF_ROOT)
    what = argv or ['bvn', 'hprob', 'pm_small', 'chain_small']
    table = {'wind': make_wind, 'bvn': make_bvn, 'hprob': make_hprob, 'pm_small': make_pm_small,
             'chain_small': make_chain_small, 'pm_full': make_pm_full, 'sprd': make_sprd, 'bayes_funcs': make_bayes_funcs,
             'kalbar_full': lambda: _make_full('kalbar'),
             'carnarvon_full': lambda: _make_full('carnarvon')}
    for w in what:
        table[w]()


if __name__ == '__main__':
    main(sys.argv[1:])
