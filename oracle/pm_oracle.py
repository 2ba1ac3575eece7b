"""CPU oracle, phase 1: per-day dispersal kernel (ParasitoidModel.py restated).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Plain numpy, fp64.

Each function cites the reference lines it follows (paths are relative to the
reference tree).  The one piece of third-party arithmetic on this path,
``scipy.stats.mvn.mvnun`` (Fortran MVNDST by A. Genz; 2-D special case
BVNMVN/BVU), is absent from the reference tree *and* from current SciPy, so it
is restated here from the published algorithm (A. Genz, "Numerical computation
of rectangular bivariate and trivariate normal and t probabilities",
Statistics and Computing 14 (2004) 251-260; TVPACK routine BVU) and checked in
tests/test_oracle.py against SciPy's compiled Genz BVU
(``scipy.special._ufuncs._bivariate_normal_cdf``) and against golden vectors
produced by the reference's own ``get_mvn_cdf_values`` / ``prob_mass``.
"""
from math import floor
import warnings

import numpy as np
from scipy import sparse
from scipy.special import erfc

# --------------------------------------------------------------------------
# Gauss-Legendre half rules used by BVU (weights, positive abscissae)
# --------------------------------------------------------------------------
_GL = {
    3: (np.array([0.1713244923791705, 0.3607615730481384, 0.4679139345726904]),
        np.array([0.9324695142031522, 0.6612093864662647, 0.2386191860831970])),
    6: (np.array([0.04717533638651177, 0.1069393259953183, 0.1600783285433464,
                  0.2031674267230659, 0.2334925365383547, 0.2491470458134029]),
        np.array([0.9815606342467191, 0.9041172563704750, 0.7699026741943050,
                  0.5873179542866171, 0.3678314989981802, 0.1252334085114692])),
    10: (np.array([0.01761400713915212, 0.04060142980038694, 0.06267204833410906,
                   0.08327674157670475, 0.1019301198172404, 0.1181945319615184,
                   0.1316886384491766, 0.1420961093183821, 0.1491729864726037,
                   0.1527533871307259]),
         np.array([0.9931285991850949, 0.9639719272779138, 0.9122344282513259,
                   0.8391169718222188, 0.7463319064601508, 0.6360536807265150,
                   0.5108670019508271, 0.3737060887154196, 0.2277858511416451,
                   0.07652652113349733])),
}

_SQRT2 = np.sqrt(2.0)
_TWOPI = 2.0 * np.pi


def phid(z):
    """Standard normal CDF (MVNDST ``MVPHI``)."""
    return 0.5 * erfc(-np.asarray(z, dtype=float) / _SQRT2)


def gl_order(r):
    """Number of Gauss-Legendre half-rule nodes BVU uses for correlation r."""
    ar = abs(r)
    return 3 if ar < 0.3 else (6 if ar < 0.75 else 10)


def bvu(dh, dk, r):
    """P(X > dh, Y > dk) for a standard bivariate normal with correlation r.

    Genz BVU; vectorised over dh, dk (broadcast), scalar r.
    """
    h, k = np.broadcast_arrays(np.asarray(dh, dtype=float),
                               np.asarray(dk, dtype=float))
    h = h.astype(float)
    k = k.astype(float)
    w, x = _GL[gl_order(r)]
    hk = h * k
    if abs(r) < 0.925:
        hs = (h * h + k * k) / 2.0
        asr = np.arcsin(r)
        acc = np.zeros_like(h)
        for wi, xi in zip(w, x):
            for s in (-1.0, 1.0):
                sn = np.sin(asr * (1.0 + s * xi) / 2.0)
                acc += wi * np.exp((sn * hk - hs) / (1.0 - sn * sn))
        return acc * asr / (2.0 * _TWOPI) + phid(-h) * phid(-k)

    # |r| >= 0.925: Drezner-Wesolowsky style expansion around |r| = 1
    if r < 0:
        k = -k
        hk = -hk
    bvn = np.zeros_like(h)
    if abs(r) < 1:
        as_ = (1.0 - r) * (1.0 + r)
        a = np.sqrt(as_)
        bs = (h - k) ** 2
        c = (4.0 - hk) / 8.0
        d = (12.0 - hk) / 16.0
        asr = -(bs / as_ + hk) / 2.0
        m = asr > -100.0
        with np.errstate(over='ignore', under='ignore', invalid='ignore'):
            t = a * np.exp(np.where(m, asr, 0.0)) * (
                1.0 - c * (bs - as_) * (1.0 - d * bs / 5.0) / 3.0
                + c * d * as_ * as_ / 5.0)
            bvn = np.where(m, t, bvn)
            m2 = -hk < 100.0
            b = np.sqrt(bs)
            t = np.exp(np.where(m2, -hk / 2.0, 0.0)) * np.sqrt(_TWOPI) * phid(-b / a) * b * (
                1.0 - c * bs * (1.0 - d * bs / 5.0) / 3.0)
            bvn = np.where(m2, bvn - t, bvn)
            a = a / 2.0
            for wi, xi in zip(w, x):
                for s in (-1.0, 1.0):
                    xs = (a * (1.0 + s * xi)) ** 2
                    rs = np.sqrt(1.0 - xs)
                    asr = -(bs / xs + hk) / 2.0
                    m = asr > -100.0
                    t = a * wi * np.exp(np.where(m, asr, 0.0)) * (
                        np.exp(np.where(m, -hk * xs / (2.0 * (1.0 + rs) ** 2), 0.0)) / rs
                        - (1.0 + c * xs * (1.0 + d * xs)))
                    bvn = np.where(m, bvn + t, bvn)
        bvn = -bvn / _TWOPI
    if r > 0:
        bvn = bvn + phid(-np.maximum(h, k))
    else:
        bvn = -bvn
        corr = np.where(h < 0, phid(k) - phid(h), phid(-h) - phid(-k))
        bvn = np.where(k > h, bvn + corr, bvn)
    return bvn


def mvn_rect(low, upp, mu, S):
    """``mvn.mvnun(low, upp, mu, S)[0]`` for 2-D finite limits.

    Follows the f2py wrapper + MVNDST/BVNMVN: standardise the limits with
    sigma = sqrt(diag S), rho = S01/(s0 s1), then the 4-term BVU difference
    (call sites: ParasitoidModel.py:340,356,366,370).  Vectorised over the
    leading axes of low/upp.
    """
    low = np.asarray(low, dtype=float)
    upp = np.asarray(upp, dtype=float)
    mu = np.asarray(mu, dtype=float)
    s0 = np.sqrt(S[0][0])
    s1 = np.sqrt(S[1][1])
    rho = S[0][1] / (s0 * s1)
    xl = (low[..., 0] - mu[0]) / s0
    xu = (upp[..., 0] - mu[0]) / s0
    yl = (low[..., 1] - mu[1]) / s1
    yu = (upp[..., 1] - mu[1]) / s1
    return bvu(xl, yl, rho) - bvu(xu, yl, rho) - bvu(xl, yu, rho) + bvu(xu, yu, rho)


# --------------------------------------------------------------------------
# Flight-probability functions
# --------------------------------------------------------------------------
def g_wind_prob(windr, aw, bw):
    """ParasitoidModel.py:231-240 -- logistic take-off scaling in wind speed."""
    return 1.0 / (1.0 + np.exp(bw * (np.asarray(windr, dtype=float) - aw)))


def f_time_prob(n, a1, b1, a2, b2):
    """ParasitoidModel.py:243-267 -- time-of-day take-off pmf on n slots."""
    t = np.linspace(0, 24 - 24.0 / n, n)
    up = 1.0 / (1.0 + np.exp(-b1 * (t - a1)))
    down = 1.0 / (1.0 + np.exp(-b2 * (t - a2)))
    lik = np.fmax(up - down, np.zeros_like(t))
    return lik / lik.sum()


def Dmat(sig_x, sig_y, rho):
    """ParasitoidModel.py:269-280 -- diffusion covariance matrix."""
    assert sig_x > 0, 'sig_x must be positive'
    assert sig_y > 0, 'sig_y must be positive'
    assert -1 <= rho <= 1, 'correlation must be between -1 and 1'
    cov = rho * sig_x * sig_y
    return np.array([[sig_x ** 2, cov], [cov, sig_y ** 2]])


def h_flight_prob(day_wind, lam, aw, bw, a1, b1, a2, b2):
    """ParasitoidModel.py:282-309 -- per-period take-off probability."""
    day_wind = np.asarray(day_wind, dtype=float)
    if day_wind.ndim > 1:
        n = day_wind.shape[0]
        windr = day_wind[:, 2]
    else:                       # single-period test form (:298-302)
        n = 1
        windr = day_wind[2]
    f = f_time_prob(n, a1, b1, a2, b2)
    g = g_wind_prob(windr, aw, bw)
    fg = f * g
    t_vec = np.linspace(1, n, n)
    tail = np.cumsum((1 - np.cumsum(f) ** 1) * (f - fg))
    return lam * (fg + fg / t_vec / np.max(f) * tail)


# --------------------------------------------------------------------------
# BVN cell masses with ring growth
# --------------------------------------------------------------------------
_RING_ORDER_CACHE = {}


def _ring_order(H):
    """(x, y) cell visiting order of ParasitoidModel.py:340-373 up to ring H."""
    if H in _RING_ORDER_CACHE:
        return _RING_ORDER_CACHE[H]
    xs, ys = [0], [0]
    for g in range(1, H + 1):
        for ii in (-g, g):
            for jj in (-g, g):
                xs.append(ii)
                ys.append(jj)
        for ii in (-g, g):
            for jj in range(-g + 1, g):
                xs.append(ii)
                ys.append(jj)
                xs.append(jj)
                ys.append(ii)
    out = (np.array(xs), np.array(ys))
    _RING_ORDER_CACHE[H] = out
    return out


def _cells_xy(cell_length, mu, S, H):
    """Cell masses indexed [x + H, y + H] for integer cell coords in [-H, H]."""
    r = cell_length / 2
    idx = np.arange(-H, H + 1)
    lo1 = idx * cell_length - r          # ParasitoidModel.py:354
    up1 = lo1 + cell_length              # :355
    low = np.stack(np.meshgrid(lo1, lo1, indexing='ij'), axis=-1)
    upp = np.stack(np.meshgrid(up1, up1, indexing='ij'), axis=-1)
    return mvn_rect(low, upp, mu, S)


def support_halfwidth(cell_length, mu, S, cdf_eps=0.001, H0=None):
    """Smallest h with 1 - (running sum in reference call order) < cdf_eps.

    Returns (h, cells_xy) with cells_xy of half-width >= h.
    """
    H = 4 if H0 is None else max(int(H0) + 1, 1)
    while True:
        cells = _cells_xy(cell_length, mu, S, H)
        ox, oy = _ring_order(H)
        running = np.cumsum(cells[ox + H, oy + H])      # sequential, call order
        ends = (2 * np.arange(H + 1) + 1) ** 2 - 1
        ok = np.nonzero(1 - running[ends] < cdf_eps)[0]
        if ok.size:
            return int(ok[0]), cells, H
        H = 2 * H + 1


def get_mvn_cdf_values(cell_length, mu, S, H0=None):
    """ParasitoidModel.py:311-380 -- (2h+1)x(2h+1) cell masses, [row, col] =
    (y = h - row, x = col - h)."""
    mu = np.asarray(mu, dtype=float)
    if mu.ndim == 0:
        mu = np.array([float(mu), float(mu)])
    h, cells, H = support_halfwidth(cell_length, mu, S, H0=H0)
    sub = cells[H - h:H + h + 1, H - h:H + h + 1]       # [x, y]
    return np.ascontiguousarray(sub.T[::-1, :])          # rows: y descending


def sprd_kernel(res, Dparams, Dlparams, sprd_factor, mean_drift=(-25., 15.)):
    """Bayes_Run.py:245-270 / Bayes_MAP.py:247-277 -- the local day-0 spread kernel the Bayes drivers prepend when
    the wind record starts a day late: the in-flow blob carried by the mean drift (integer cells + remainder, Python
    floor division / modulo), mixed with the out-of-flow blob, centre topped up to unit mass.  Dense (mlen, mlen)."""
    xi = int(mean_drift[0] // res)
    xr = mean_drift[0] % res
    yi = int(mean_drift[1] // res)
    yr = mean_drift[1] % res
    longsprd = get_mvn_cdf_values(res, np.array([xr, yr]), Dmat(*Dparams))
    shrtsprd = get_mvn_cdf_values(res, np.array([0., 0.]), Dmat(*Dlparams))
    mlen = int(max(longsprd.shape[0], shrtsprd.shape[0]) + max(abs(xi), abs(yi)) * 2)
    sprd = np.zeros((mlen, mlen))
    lo, hi = mlen // 2 - longsprd.shape[0] // 2, mlen // 2 + longsprd.shape[0] // 2 + 1
    sprd[lo - yi:hi - yi, lo + xi:hi + xi] = longsprd * sprd_factor
    lo, hi = mlen // 2 - shrtsprd.shape[0] // 2, mlen // 2 + shrtsprd.shape[0] // 2 + 1
    sprd[lo:hi, lo:hi] += shrtsprd * (1 - sprd_factor)
    sprd[mlen // 2, mlen // 2] += max(0, 1 - sprd.sum())
    return sprd


# --------------------------------------------------------------------------
# Wind input (ParasitoidModel.py:64-227): the path's input producer
# --------------------------------------------------------------------------
def read_wind_file(site_name):
    """ParasitoidModel.py:64-132 -- raw wind series {day: ndarray(times, 3)} of (windx, windy, windr) and the sorted
    day list; components below 1e-4 are zeroed."""
    rows = {}
    with open(site_name + 'wind.txt') as fobj:
        for line in fobj:
            cols = line.split()
            if not cols:
                continue
            wx, wy = float(cols[1]), float(cols[2])
            if abs(wx) < 10e-5:
                wx = 0
            if abs(wy) < 10e-5:
                wy = 0
            wr = np.sqrt(wx ** 2 + wy ** 2)
            if abs(wr) < 10e-5:
                wr = 0
            rows.setdefault(int(cols[0]), []).append((wx, wy, wr))
    wind = {day: np.array(v, dtype=float) for day, v in rows.items()}
    return wind, sorted(wind)


def get_wind_data(site_name, interp_num, start_time):
    """Linearly interpolated wind, ``interp_num`` points per raw interval
    (ParasitoidModel.py:136-227).  '00:00': the day's last interval interpolates towards the
    next day's first sample (the final day repeats its last sample); '00:30':
    the day's first interval interpolates from the previous day's last sample
    (the first day repeats its first sample)."""
    raw, days = read_wind_file(site_name)
    if start_time not in ('00:00', '00:30'):
        raise ValueError("start_time must be either '00:00' or '00:30'")
    npts = raw[days[0]].shape[0]
    w1 = np.linspace(0, 1, interp_num + 1)[:-1][:, None]       # weight of the later sample
    w0 = 1 - w1

    def blend(a, b):
        return w0 * a + w1 * b

    wind = {}
    for n, day in enumerate(days):
        r = raw[day]
        out = np.zeros((npts * interp_num, 3))
        first = 0 if start_time == '00:00' else 1
        for k in range(npts - 1):
            out[(k + first) * interp_num:(k + first + 1) * interp_num] = blend(r[k], r[k + 1])
        if start_time == '00:00':
            if n + 1 < len(days):
                out[(npts - 1) * interp_num:] = blend(r[-1], raw[day + 1][0])
                out[:, 2] = np.sqrt(out[:, 0] ** 2 + out[:, 1] ** 2)
            else:
                out[:, 2] = np.sqrt(out[:, 0] ** 2 + out[:, 1] ** 2)
                out[(npts - 1) * interp_num:] = r[-1]
        else:
            if n == 0:
                out[:interp_num] = r[0]
            else:
                out[:interp_num] = blend(raw[day - 1][-1], r[0])
            out[:, 2] = np.sqrt(out[:, 0] ** 2 + out[:, 1] ** 2)
        wind[day] = out
    return wind, days


# --------------------------------------------------------------------------
# Threshold / renormalise (CalcSol.py:112-136), needed by prob_mass
# --------------------------------------------------------------------------
def r_small_vals(A, prob_model=False, negval=1e-8):
    """CalcSol.py:112-136 -- drop entries < negval; if prob_model spread the
    missing mass uniformly over the survivors so the result sums to 1."""
    if not sparse.isspmatrix_coo(A):
        A = sparse.coo_matrix(A)
    keep = ~(A.data < negval)
    data = A.data[keep]
    out = sparse.coo_matrix((data, (A.row[keep], A.col[keep])), A.shape)
    if prob_model:
        out.data += (1 - out.data.sum()) / out.data.size
    return out


# --------------------------------------------------------------------------
# Per-day kernel
# --------------------------------------------------------------------------
def drift_for_period(day, wind_data, t_indx, n_periods, mu_r, single):
    """Flight-averaged advection for one take-off period in metres
    (ParasitoidModel.py:439-472)."""
    day_wind = wind_data[day]
    if single:
        mu_v = np.array(day_wind[0:2], dtype=float)
        periods = 1
    else:
        periods = day_wind.shape[0]
        if n_periods > 1:
            if t_indx + n_periods - 1 < periods:
                mu_v = np.sum(day_wind[t_indx:t_indx + n_periods, 0:2], 0) / n_periods
            elif day + 1 in wind_data:
                if t_indx != periods - 1:
                    mu_v = np.sum(day_wind[t_indx:, 0:2], 0)
                else:
                    mu_v = np.array(day_wind[-1, 0:2])
                wrap = n_periods - (periods - t_indx)
                if wrap != 1:
                    mu_v += np.sum(wind_data[day + 1][:wrap, 0:2], 0)
                else:
                    mu_v += wind_data[day + 1][0, 0:2]
                mu_v /= n_periods
            else:
                if t_indx != periods - 1:
                    mu_v = np.sum(day_wind[t_indx:, 0:2], 0) / (periods - t_indx)
                else:
                    mu_v = np.array(day_wind[-1, 0:2])
        else:
            mu_v = np.array(day_wind[t_indx, 0:2])
    mu_v = mu_v * (3600 * 24 * (n_periods / periods))
    mu_v = mu_v * mu_r
    return mu_v


def prob_mass(day, wind_data, hparams, Dparams, Dlparams, mu_r, n_periods,
              rad_dist, rad_res, start_time=None, details=None):
    """ParasitoidModel.py:384-613 -- one day's displacement pmf as COO.

    ``details`` (optional dict) receives the pre-threshold dense grid and the
    bookkeeping scalars the parity tests compare on: 'pmf_pre' (dense,
    dom_len^2, after the local-diffusion blob, before r_small_vals), 'loss',
    'total_flight_prob', 'hprob', 'offsets' (per period: row_cent, col_cent,
    h), 'warned'.
    """
    dom_len = rad_res * 2 + 1
    cell = rad_dist / rad_res
    pmf = np.zeros((dom_len, dom_len))
    day_wind = wind_data[day]
    hprob = h_flight_prob(day_wind, *hparams)
    S = Dmat(*Dparams)
    Sl = Dmat(*Dlparams)
    loss = 0.0
    single = not (np.ndim(day_wind) > 1)
    periods = 1 if single else day_wind.shape[0]
    start_indx = 0 if start_time is None else floor(start_time * periods)
    warned = False
    offsets = []
    h_guess = None
    hp = np.atleast_1d(hprob)
    for t in range(start_indx, periods):
        mu_v = drift_for_period(day, wind_data, t, n_periods, mu_r, single)
        shift = np.round(mu_v / cell)
        cdf_mu = mu_v - shift * cell
        cdf_mat = get_mvn_cdf_values(cell, cdf_mu, S, H0=h_guess)
        nr = cdf_mat.shape[0] // 2
        h_guess = nr
        col_c = rad_res + int(np.round(mu_v[0] / cell))
        row_c = rad_res + int(np.round(-mu_v[1] / cell))
        offsets.append((row_c, col_c, nr))
        r0, r1 = row_c - nr, row_c + nr          # inclusive window in pmf
        c0, c1 = col_c - nr, col_c + nr
        ks, ke = 0, cdf_mat.shape[0]             # window in cdf_mat (rows)
        ls, le = 0, cdf_mat.shape[1]
        if r1 + 1 > dom_len:
            ke = max(0, ke - (r1 + 1 - dom_len))
            r1 = dom_len - 1
        if c1 + 1 > dom_len:
            le = max(0, le - (c1 + 1 - dom_len))
            c1 = dom_len - 1
        if r0 < 0:
            ks = max(ks - r0, 0)
            r0 = 0
        if c0 < 0:
            ls = max(ls - c0, 0)
            c0 = 0
        if not (-1e-9 <= hp[t] <= 1.000000001):
            raise AssertionError('hprob out of bounds at t_indx {}'.format(t),
                                 'hprob[t_indx]={}'.format(hp[t]))
        try:
            # numpy basic slicing semantics are part of the reference's
            # behaviour here (negative stops wrap) -- keep them (:539-558).
            pmf[r0:r1 + 1, c0:c1 + 1] += hp[t] * cdf_mat[ks:ke, ls:le]
            if ks > 0 or ke < cdf_mat.shape[0] or ls > 0 or le < cdf_mat.shape[1]:
                loss += (1 - cdf_mat[ks:ke, ls:le].sum()) * hp[t]
        except ValueError:
            if not warned:
                warnings.warn('Index error in calculating prob_mass.\n'
                              'Day: {}, Period: {}, mu_v: {}\n'.format(day, t, mu_v) +
                              'Wind advection during this period appears to be greater'
                              ' than the size of the domain.\n'
                              'Wasps flying during this time will be considered lost.',
                              RuntimeWarning)
                warned = True
            loss += hp[t]

    pmfsum = pmf.sum()
    total = pmfsum + loss
    assert loss >= 0.0, 'negative loss'
    assert pmf.min() >= -1e-8, 'pmf.min() less than zero, first block'
    assert pmfsum <= 1.00001, 'flight prob > 1, first block'
    if total < 0.99999:
        blob = get_mvn_cdf_values(cell, np.array([0., 0.]), Sl)
        nr = blob.shape[0] // 2
        pmf[rad_res - nr:rad_res + nr + 1, rad_res - nr:rad_res + nr + 1] += (1 - total) * blob
        total2 = pmf.sum() + loss
        assert pmf.min() >= -1e-8, 'pmf.min() less than zero'
        assert total2 <= 1.00001, 'flight prob > 1'
    if details is not None:
        details['pmf_pre'] = pmf.copy()
        details['loss'] = loss
        details['total_flight_prob'] = total
        details['hprob'] = hp.copy()
        details['offsets'] = np.array(offsets, dtype=np.int64).reshape(-1, 3)
        details['warned'] = warned
    coo = r_small_vals(sparse.coo_matrix(pmf), prob_model=True)
    I, J, V = coo.row, coo.col, coo.data
    rad = int(max(np.fabs(I - rad_res).max(), np.fabs(J - rad_res).max()))
    return sparse.coo_matrix((V, (I - rad_res + rad, J - rad_res + rad)),
                             shape=(rad * 2 + 1, rad * 2 + 1))
