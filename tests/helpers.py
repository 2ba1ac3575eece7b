"""Shared helpers for the parity tests (fixtures, metrics)."""
import os

import numpy as np
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')

# Run.py:68-83 defaults
HPARAMS = (1., 1.263, 3.913, 7.302, 2.614, 23.999, 2.350)
DPARAMS = (171.82, 144.58, 0.253)
DLPARAMS = (7.096, 7.260, 0.000)
MU_R = 1.179

# parity bars of the north star
MAX_ABS = 1e-10
REL_L1 = 1e-9
MASS = 1e-12

SITES = {'kalbar': '00:00', 'carnarvon': '00:30'}


def load(name):
    return np.load(os.path.join(GOLD, name + '.npz'))


def coo(z, prefix):
    shp = tuple(int(s) for s in z[prefix + '_shape'])
    return sparse.coo_matrix((z[prefix + '_val'], (z[prefix + '_row'], z[prefix + '_col'])), shape=shp)


def write_wind_file(tmpdir, site):
    """Recreate data/<site>wind.txt from the golden raw series; returns the
    site_name prefix to hand to get_wind_data."""
    z = load('wind')
    days, raw = z[site + '_days'], z[site + '_raw']
    prefix = os.path.join(str(tmpdir), site)
    with open(prefix + 'wind.txt', 'w') as fobj:
        for d, block in zip(days, raw):
            for wx, wy, _ in block:
                fobj.write('%d\t%.17g\t%.17g\n' % (d, wx, wy))
    return prefix


def rel_l1(a, b):
    den = np.abs(b).sum()
    return np.abs(a - b).sum() / den if den > 0 else np.abs(a - b).sum()


def assert_parity(got, ref, what='', max_abs=MAX_ABS, l1=REL_L1):
    got = np.asarray(got, dtype=float)
    ref = np.asarray(ref, dtype=float)
    assert got.shape == ref.shape, '{}: shape {} vs {}'.format(what, got.shape, ref.shape)
    scale = max(1.0, np.abs(ref).max())
    ma = np.abs(got - ref).max()
    assert ma <= max_abs * scale, '{}: max-abs {:.3e} > {:.1e}'.format(what, ma, max_abs * scale)
    r = rel_l1(got, ref)
    assert r <= l1, '{}: relative L1 {:.3e} > {:.1e}'.format(what, r, l1)


def assert_thresholded_parity(got, ref, negval=1e-8, what='', tol=1e-12, max_abs=MAX_ABS, l1=REL_L1):
    """Compare two thresholded (r_small_vals) grids: cells present in only one
    of them must sit within `tol` (relative to negval scale) of the drop
    threshold -- a keep/drop flip of a borderline cell -- and
    the common support must agree to the parity bars."""
    got = np.asarray(got, dtype=float)
    ref = np.asarray(ref, dtype=float)
    assert got.shape == ref.shape, '{}: shape {} vs {}'.format(what, got.shape, ref.shape)
    only = (got != 0) != (ref != 0)
    if only.any():
        vals = np.where(got != 0, got, ref)[only]
        assert np.all(np.abs(vals - negval) < 1e-4 * negval), \
            '{}: support differs away from the threshold ({} cells)'.format(what, int(only.sum()))
        assert only.sum() <= 3, '{}: {} borderline flips'.format(what, int(only.sum()))
    both = ~only
    assert_parity(np.where(both, got, 0.0), np.where(both, ref, 0.0), what, max_abs, l1)


def recentre(pmf, rad_res):
    """Run.py:454-458."""
    pmf = sparse.coo_matrix(pmf)
    off = rad_res - pmf.shape[0] // 2
    D = 2 * rad_res + 1
    return sparse.coo_matrix((pmf.data, (pmf.row + off, pmf.col + off)), shape=(D, D))
