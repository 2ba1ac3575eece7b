"""Drop-in for the reference's ``ParasitoidModel`` module (model physics and
per-day dispersal-kernel construction), computed on the GPU.

Same function names, argument meaning, return types and error behaviour as
``ParasitoidModel.py`` in mountaindust/Parasitoids (line numbers below refer
to that file).  The arithmetic -- flight probability, flight-averaged drift,
Genz bivariate-normal cell masses, accumulation, threshold/renormalise --
runs in libpkb200's sm_100a kernels in fp64; this module only marshals
arguments and rebuilds the scipy.sparse return values.  The wind/emergence
file readers are host-side input producers ("not under MCMC", :142) and stay
in Python.
"""
import ctypes as C
import warnings

import numpy as np
from scipy import sparse

from . import _abi
from . import _lib


# ---------------------------------------------------------------------------
# input producers (host side)
# ---------------------------------------------------------------------------
def emergence_data(site_name):
    """Observed emergence counts ``em[field][date]`` (:28-60)."""
    em = {}
    with open(site_name + 'emergence.txt', 'r') as fobj:
        fields = fobj.readline().split()[1:]
        for name in fields:
            em[name] = {}
        for line in fobj:
            cols = line.split()
            if not cols:
                continue
            date = int(cols[0])
            for name, val in zip(fields, cols[1:]):
                em[name][date] = int(val)
    return em


def read_wind_file(site_name):
    """Raw wind series ``{day: ndarray(times, 3)}`` of (windx, windy, windr)
    plus the sorted day list (:64-132).  Components below 1e-4 are zeroed."""
    rows = {}
    with open(site_name + 'wind.txt') as fobj:
        for line in fobj:
            cols = line.split()
            if not cols:
                continue
            day = int(cols[0])
            wx = float(cols[1])
            wy = float(cols[2])
            if abs(wx) < 10e-5:
                wx = 0
            if abs(wy) < 10e-5:
                wy = 0
            wr = np.sqrt(wx ** 2 + wy ** 2)
            if abs(wr) < 10e-5:
                wr = 0
            rows.setdefault(day, []).append((wx, wy, wr))
    wind = {day: np.array(v, dtype=float) for day, v in rows.items()}
    return wind, sorted(wind)


def get_wind_data(site_name, interp_num, start_time):
    """Linearly interpolated wind, ``interp_num`` points per raw interval
    (:136-227).  '00:00': the day's last interval interpolates towards the
    next day's first sample (the final day repeats its last sample); '00:30':
    the day's first interval interpolates from the previous day's last sample
    (the first day repeats its first sample).  The file is parsed on the host;
    the interpolation runs on the device (k_wind_interp) with numpy's own
    weights and operation order, so the arrays equal the reference's bit for bit."""
    raw, days = read_wind_file(site_name)
    if start_time not in ('00:00', '00:30'):
        raise ValueError("start_time must be either '00:00' or '00:30'")
    if days != list(range(days[0], days[0] + len(days))):
        raise KeyError('wind days must be consecutive integers (:178,:214)')
    npts = raw[days[0]].shape[0]
    stacked = _lib.as_f64(np.stack([raw[d] for d in days]))
    out = np.empty((len(days), npts * int(interp_num), 3))
    _lib.check(_lib.lib().pkb_wind_interp(_lib.ctx().h, _lib.dptr(stacked), len(days), npts, int(interp_num),
                                          0 if start_time == '00:00' else 1, _lib.dptr(out)))
    return {day: out[n] for n, day in enumerate(days)}, days


# ---------------------------------------------------------------------------
# flight probability (device: k_hprob)
# ---------------------------------------------------------------------------
def _hprob_call(wind, single, hparams, want):
    wind = _lib.as_f64(wind)
    periods = 1 if single else wind.shape[0]
    hp = (C.c_double * 7)(*[float(v) for v in hparams])
    h = np.empty(periods)
    f = np.empty(periods)
    g = np.empty(periods)
    _lib.check(_lib.lib().pkb_hprob(_lib.ctx().h, _lib.dptr(wind), periods, 1 if single else 0, hp,
                                    _lib.dptr(h), _lib.dptr(f), _lib.dptr(g)))
    return {'h': h, 'f': f, 'g': g}[want]


def g_wind_prob(windr, aw, bw):
    """Logistic take-off scaling in wind speed (:231-240)."""
    wr = np.asarray(windr, dtype=float)
    flat = np.atleast_1d(wr).ravel()
    wind = np.zeros((flat.size, 3))
    wind[:, 2] = flat
    g = _hprob_call(wind, False, (1., aw, bw, 6., 1., 18., 1.), 'g')
    return g.reshape(wr.shape) if wr.ndim else float(g[0])


def f_time_prob(n, a1, b1, a2, b2):
    """Time-of-day take-off pmf on ``n`` slots (:243-267)."""
    n = int(n)
    return _hprob_call(np.zeros((n, 3)), False, (1., 0., 1., a1, b1, a2, b2), 'f')


def Dmat(sig_x, sig_y, rho):
    """Diffusion covariance matrix (:269-280)."""
    assert sig_x > 0, 'sig_x must be positive'
    assert sig_y > 0, 'sig_y must be positive'
    assert -1 <= rho <= 1, 'correlation must be between -1 and 1'
    return np.array([[sig_x ** 2, rho * sig_x * sig_y], [rho * sig_x * sig_y, sig_y ** 2]])


def h_flight_prob(day_wind, lam, aw, bw, a1, b1, a2, b2):
    """Probability density of flying in each period of the day (:282-309)."""
    day_wind = np.asarray(day_wind, dtype=float)
    single = day_wind.ndim == 1
    h = _hprob_call(day_wind, single, (lam, aw, bw, a1, b1, a2, b2), 'h')
    return h if not single else h[:1]


# ---------------------------------------------------------------------------
# BVN cell masses (device: k_mvn_cdf)
# ---------------------------------------------------------------------------
def get_mvn_cdf_values(cell_length, mu, S):
    """Cell masses of N(mu, S) on the smallest (2h+1)^2 lattice holding all
    but 0.001 of the mass; ``[row, col] = (y = h - row, x = col - h)`` (:311-380)."""
    mu = np.asarray(mu, dtype=float)
    if mu.ndim == 0:
        mu = np.array([float(mu), float(mu)])
    S = np.asarray(S, dtype=float)
    cov = (C.c_double * 3)(S[0, 0], S[1, 1], S[0, 1])
    mu_c = (C.c_double * 2)(mu[0], mu[1])
    cap = 101 * 101
    while True:
        out = np.empty(cap)
        h = C.c_int(-1)
        rc = _lib.lib().pkb_mvn_cdf(_lib.ctx().h, float(cell_length), mu_c, cov, _lib.dptr(out), cap, C.byref(h))
        if rc == _abi.PKB_ELIMIT and cap < 8193 * 8193:
            cap *= 16
            continue
        _lib.check(rc)
        n = 2 * h.value + 1
        return out[:n * n].reshape(n, n).copy()


# ---------------------------------------------------------------------------
# per-day kernel (device: k_drift / k_period / k_day_finalize)
# ---------------------------------------------------------------------------
def _day_args(hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res, start_time, wind_day, single):
    a = _abi.DayArgs()
    a.hparams[:] = [float(v) for v in hparams]
    a.dparams[:] = [float(v) for v in Dparams]
    a.dlparams[:] = [float(v) for v in Dlparams]
    a.mu_r = float(mu_r)
    a.rad_dist = float(rad_dist)
    a.start_time = -1.0 if start_time is None else float(start_time)
    a.n_periods = int(n_periods)
    a.rad_res = int(rad_res)
    a.wind_day = int(wind_day)
    a.single = 1 if single else 0
    return a


_ASSERTS = (
    (_abi.ST_HPROB_RANGE, 'hprob out of bounds'),                 # :528-537
    (_abi.ST_NEG_LOSS, 'negative loss'),                          # :568-580
    (_abi.ST_PMF_NEG, 'pmf.min() less than zero, first block'),
    (_abi.ST_PMF_GT1, 'flight prob > 1, first block'),
    (_abi.ST_PMF_NEG2, 'pmf.min() less than zero'),               # :588-599
    (_abi.ST_TOT_GT1, 'flight prob > 1'),
)


def _raise_for_status(meta, day, params):
    for bit, msg in _ASSERTS:
        if meta.status & bit:
            raise AssertionError(msg, 'day = {}'.format(day), 'loss = {}'.format(meta.loss),
                                 'pmf.sum() = {}'.format(meta.pmfsum), 'params = {}'.format(params))
    if meta.status & _abi.ST_WARNED:
        warnings.warn('Index error in calculating prob_mass.\nDay: {}\n'
                      'Wind advection during some period appears to be greater than the size of the domain.\n'
                      'Wasps flying during this time will be considered lost.'.format(day), RuntimeWarning)


class KernelSet(object):
    """Device-resident per-day kernels of one ``pkb_kernels_build`` call."""

    def __init__(self, handle, nprob):
        self.h = handle
        self.nprob = nprob

    def meta(self, i):
        m = _abi.DayMeta()
        _lib.check(_lib.lib().pkb_kset_meta(self.h, i, C.byref(m)))
        return m

    def dense(self, i):
        n = 2 * self.meta(i).rad + 1
        out = np.empty((n, n))
        _lib.check(_lib.lib().pkb_kset_get(self.h, i, _lib.dptr(out)))
        return out

    def pre(self, i):
        """Pre-threshold accumulation window (needs keep_pre)."""
        n = 2 * _lib.lib().pkb_kset_racc(self.h) + 1
        out = np.empty((n, n))
        _lib.check(_lib.lib().pkb_kset_get_pre(self.h, i, _lib.dptr(out)))
        return out

    def periods(self, i, periods):
        rch = np.empty((periods, 3), dtype=np.int32)
        hp = np.empty(periods)
        _lib.check(_lib.lib().pkb_kset_periods(self.h, i, _lib.iptr(rch), _lib.dptr(hp)))
        return rch, hp

    def close(self):
        if self.h:
            _lib.lib().pkb_kset_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_kernels(wind, args, keep_pre=False):
    """wind: ndarray (nd_wind, periods, 3); args: list of _abi.DayArgs."""
    wind = _lib.as_f64(wind)
    arr = (_abi.DayArgs * len(args))(*args)
    h = C.c_void_p()
    _lib.check(_lib.lib().pkb_kernels_build(_lib.ctx().h, _lib.dptr(wind), wind.shape[0], wind.shape[1], arr,
                                            len(args), 1 if keep_pre else 0, C.byref(h)))
    return KernelSet(h, len(args))


def sprd_kernel(res, Dparams, Dlparams, sprd_factor, mean_drift=(-25., 15.)):
    """The local day-0 spread kernel the Bayes drivers build inline when the wind record starts a day late
    (Bayes_Run.py:245-270, Bayes_MAP.py:247-277), as a dense (mlen, mlen) array: ``sprd_factor`` times the in-flow
    blob ``get_mvn_cdf_values(res, drift remainder, Dmat(*Dparams))`` shifted by the whole cells of ``mean_drift``,
    plus ``1 - sprd_factor`` times the out-of-flow blob, centre topped up to unit mass.  ``Run.solve(...,
    sprd_factor=...)`` builds and uses it on the device; this function hands it back for inspection."""
    a = _abi.DayArgs()
    a.dparams[:] = [float(v) for v in Dparams]
    a.dlparams[:] = [float(v) for v in Dlparams]
    a.rad_res = 4096                      # only the cell size matters: rad_dist / rad_res = res
    a.rad_dist = float(res) * a.rad_res
    a.start_time = -1.0
    a.n_periods = 1
    a.kind = 1
    a.sprd_factor = float(sprd_factor)
    a.sprd_drift[:] = [float(mean_drift[0]), float(mean_drift[1])]
    ks = build_kernels(np.zeros((1, 1, 3)), [a])
    try:
        return ks.dense(0)
    finally:
        ks.close()


def _stack_wind(day, wind_data):
    """(wind array, single) for one day plus its successor if present (:449)."""
    dw = np.asarray(wind_data[day], dtype=float)
    if dw.ndim == 1:
        return dw.reshape(1, 1, 3), True
    if day + 1 in wind_data:
        return np.stack([dw, np.asarray(wind_data[day + 1], dtype=float)]), False
    return dw[None], False


def prob_mass(day, wind_data, hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res,
              start_time=None, details=None):
    """One day's displacement pmf as a ``scipy.sparse.coo_matrix`` (:384-613).

    ``details`` (optional dict, not in the reference) receives the dense
    pre-threshold window and bookkeeping scalars for parity checks."""
    wind, single = _stack_wind(day, wind_data)
    a = _day_args(hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res, start_time, 0, single)
    ks = build_kernels(wind, [a], keep_pre=details is not None)
    try:
        meta = ks.meta(0)
        _raise_for_status(meta, day, (hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res))
        dense = ks.dense(0)
        if details is not None:
            details['pre_window'] = ks.pre(0)
            details['loss'] = meta.loss
            details['total_flight_prob'] = meta.total
            details['rad'] = meta.rad
            details['status'] = meta.status
            rch, hp = ks.periods(0, wind.shape[1])
            details['offsets'] = rch
            details['hprob'] = hp
    finally:
        ks.close()
    return sparse.coo_matrix(dense)


def prob_mass_batch(pm_args):
    """``Pool().starmap(prob_mass, pm_args)`` (Run.py:415-425, Bayes_Run.py:236-272)
    as ONE device batch.  Every tuple is a ``prob_mass`` argument list; all
    must share ``wind_data``, ``rad_res``.  Returns the list of COO matrices."""
    if not pm_args:
        return []
    wind_data = pm_args[0][1]
    days = sorted(wind_data)
    row = {d: i for i, d in enumerate(days)}
    if np.ndim(wind_data[days[0]]) == 1:
        return [prob_mass(*a) for a in pm_args]
    if days != list(range(days[0], days[0] + len(days))):
        return [prob_mass(*a) for a in pm_args]
    wind = np.stack([np.asarray(wind_data[d], dtype=float) for d in days])
    args = []
    for a in pm_args:
        start = a[9] if len(a) > 9 else None
        args.append(_day_args(a[2], a[3], a[4], a[5], a[6], a[7], a[8], start, row[a[0]], False))
    ks = build_kernels(wind, args)
    try:
        out = []
        for i, a in enumerate(pm_args):
            _raise_for_status(ks.meta(i), a[0], a[2:9])
            out.append(sparse.coo_matrix(ks.dense(i)))
    finally:
        ks.close()
    return out
