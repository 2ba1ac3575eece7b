"""Diagnostic: per-day spectral-step / pad-content record of a small calm-wind solve (GPU)."""
import sys, warnings, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import helpers as H
from parasitoids_b200 import Run, _lib
rng = np.random.default_rng(5)
nd, periods, rad_res, rad_dist = 10, 96, 140, 7000.0
w = np.zeros((nd, periods, 3))
for c in range(2):
    x = np.cumsum(rng.normal(0, 0.05, nd * periods)).reshape(nd, periods)
    w[:, :, c] = 0.15 * np.sin(np.linspace(0, 6, nd * periods)).reshape(nd, periods) + x * 0.1
w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, rad_dist, rad_res)
ctx = _lib.ctx()
for windows in (0, 1):
    ctx.set_option('windows', windows)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        res = Run.solve(w, nd, *args, want_coo=False, want_dense=True)
    print('windows', windows, 'radii', res.radii(), 'P', res.P, 'N', res.N, 'winsteps', res.window_steps())
    print(' spec', res.spectral_steps(), 'flags', res.flags())
    print(' padabs', ['%.1e' % res.day_meta(d)[1].padabs for d in range(nd)])
    print(' padmax', ['%.1e' % res.day_meta(d)[1].padmax for d in range(nd)])
    res.close()
