"""Drop-in for the hot-path half of the reference's ``Run`` module: ``Params``
(Run.py:34-384) and ``main(params)`` (Run.py:388-520) without plotting.

``main`` runs the whole forward solve -- per-day kernels and the convolution
chain -- in ONE library call (``pkb_solve``): the kernels never leave the GPU
between phase 1 and phase 2 and only the thresholded COO triplets come back.
``solve`` is the same call with the knobs the benchmark and the batched
likelihood path need.
"""
import ctypes as C
import json
import os
import time

import numpy as np
from scipy import sparse

from . import _abi
from . import _lib
from . import ParasitoidModel as PM


class Params(object):
    """Parameters of a model run; same attribute names, defaults and dataset
    presets as the reference (Run.py:41-157).  ``config.txt`` is read if it
    exists but never created."""
    OUTPUT = True
    PLOT = False      # plotting is out of scope here (Plot_Result needs matplotlib)
    CUDA = True

    def __init__(self):
        self.PROB_MODEL = True
        self.dataset = 'kalbar'
        self.my_datasets()
        self.domain_info = (10000.0, 400)
        self.interp_num = 30
        self.ndays = -1
        self.g_params = (1.263, 3.913)
        self.f_params = (7.302, 2.614, 23.999, 2.350)
        self.Dparams = (171.82, 144.58, 0.253)
        self.Dlparams = (7.096, 7.260, 0.000)
        self.lam = 1.
        self.mu_r = 1.179
        self.n_periods = 30
        self.maps_key = None
        self.maps_service = 'Google'
        self.min_ndays = 6
        self.default_chg()

    def my_datasets(self):
        if self.dataset is None:
            self.site_name = 'data/carnarvonearl'
            self.start_time = '00:30'
            self.coord = None
            self.r_dur = None
            self.r_dist = None
            self.r_start = None
            self.r_number = None
        elif self.dataset == 'carnarvon':
            self.site_name = 'data/carnarvonearl'
            self.start_time = '00:30'
            self.coord = (-24.851614, 113.731267)
            self.r_dur = 5
            self.r_dist = 'uniform'
            self.r_start = 0.354
            self.r_number = 40000
        elif self.dataset == 'kalbar':
            self.site_name = 'data/kalbar'
            self.start_time = '00:00'
            self.coord = (-27.947131, 152.584171)
            self.r_dur = 1
            self.r_dist = 'uniform'
            self.r_start = None
            self.r_number = 130000
        else:
            print('Unknown dataset in Params.dataset.')
        stamp = time.strftime('%m%d-%H%M')
        base = self.dataset if self.dataset is not None else ''
        if self.PROB_MODEL:
            self.outfile = 'output/' + base + stamp
        else:
            self.outfile = 'output/' + (base + '_pop' if base else 'poprun') + stamp

    # ---- emergence distributions (Run.py:159-183) ---------------------------
    def uniform(self, day):
        return 1. / self.r_dur

    def custom(self, day):
        pass

    def r_mthd(self):
        if self.r_dist == 'uniform':
            return self.uniform
        elif self.r_dist == 'custom':
            return self.custom

    # ---- parameter changes (Run.py:185-352) ---------------------------------
    def default_chg(self):
        try:
            with open('config.txt', 'r') as fobj:
                for line in fobj:
                    line = line.split('#', 1)[0]
                    words = line.split('=')
                    if len(words) > 1:
                        self.chg_param(words[0].strip(), words[1].strip())
            self.my_datasets()
        except FileNotFoundError:
            pass

    def cmd_line_chg(self, args):
        """``--flag`` and ``key=value`` arguments (Run.py:218-260)."""
        for argstr in args:
            if argstr[0:2] == '--':
                flag = argstr[2:].lower()
                if flag == 'kalbar':
                    self.dataset = 'kalbar'
                    self.my_datasets()
                elif flag == 'carnarvon':
                    self.dataset = 'carnarvon'
                    self.my_datasets()
                elif flag == 'prob':
                    self.PROB_MODEL = True
                    self.my_datasets()
                elif flag == 'pop':
                    self.PROB_MODEL = False
                    self.my_datasets()
                elif flag == 'no_output':
                    self.OUTPUT = False
                elif flag == 'output':
                    self.OUTPUT = True
                elif flag in ('no_plot', 'plot', 'no_cuda', 'cuda'):
                    pass    # plotting is out of scope; the device path is the only path
                else:
                    raise ValueError('Unrecognized option {0}.'.format(argstr))
            else:
                arg, eq, val = argstr.partition('=')
                try:
                    self.chg_param(arg, val)
                except Exception:
                    print('Unrecognized parameter pair {0}.'.format(argstr))
                    raise

    def chg_param(self, arg, val):
        """Change one parameter given as strings (Run.py:264-352)."""
        if arg == 'outfile':
            self.outfile = val
        elif arg == 'dataset':
            self.dataset = None if val == 'None' else val
            self.my_datasets()
        elif arg == 'site_name':
            self.site_name = val
        elif arg == 'start_time':
            self.start_time = val
        elif arg == 'r_dist':
            self.r_dist = val
        elif arg in ('domain_info', 'g_params', 'f_params', 'Dparams', 'Dlparams', 'coord'):
            vals = [v for v in val.strip('()[] ').split(',') if v.strip()]
            if arg == 'domain_info':
                self.domain_info = (float(vals[0]), int(vals[1]))
            else:
                setattr(self, arg, tuple(float(v) for v in vals))
        elif arg in ('interp_num', 'ndays', 'n_periods', 'min_ndays', 'r_dur'):
            setattr(self, arg, int(val))
        elif arg in ('lam', 'mu_r', 'r_number'):
            setattr(self, arg, float(val))
        elif arg == 'r_start':
            self.r_start = None if val == 'None' else float(val)
        elif arg in ('maps_key', 'maps_service'):
            setattr(self, arg, val)
        elif arg in ('OUTPUT', 'PLOT', 'CUDA', 'PROB_MODEL', 'output', 'plot', 'cuda', 'prob_model'):
            setattr(self, arg.upper(), val == 'True')
        else:
            raise ValueError('Parameter {0} not found.'.format(arg))

    def file_read_chg(self, filename):
        """Read parameters saved by ``main`` (Run.py:355-368)."""
        if filename.rstrip()[-5:] != '.json':
            filename += '.json'
        with open(filename) as fobj:
            param_dict = json.load(fobj)
        for key in param_dict:
            if isinstance(param_dict[key], list):
                param_dict[key] = tuple(param_dict[key])
        self.__dict__.update(param_dict)

    # ---- argument marshalling (Run.py:374-384) ------------------------------
    def get_model_params(self):
        hparams = (self.lam, *self.g_params, *self.f_params)
        return (hparams, self.Dparams, self.Dlparams, self.mu_r, self.n_periods, *self.domain_info)

    def get_wind_params(self):
        return (self.site_name, self.interp_num, self.start_time)


# ---------------------------------------------------------------------------
class SolveResult(object):
    """Outputs of one fused solve (wraps a ``pkb_result``)."""

    def __init__(self, handle):
        self.h = handle
        nd, D, P, N, ms = (C.c_int() for _ in range(5))
        _lib.check(_lib.lib().pkb_result_info(handle, *(C.byref(v) for v in (nd, D, P, N, ms))))
        self.ndays, self.dom_len, self.P, self.N, self.max_shape = nd.value, D.value, P.value, N.value, ms.value

    def day_meta(self, day):
        km, sm = _abi.DayMeta(), _abi.StepMeta()
        _lib.check(_lib.lib().pkb_result_day_meta(self.h, day, C.byref(km), C.byref(sm)))
        return km, sm

    def window_steps(self):
        """Chain steps that ran on a support-window torus (see ChainDims::win in csrc/chain.cuh)."""
        n = C.c_int()
        _lib.check(_lib.lib().pkb_result_window_steps(self.h, C.byref(n)))
        return n.value

    def flags(self):
        return [bool(self.day_meta(d)[1].flag) for d in range(self.ndays)]

    def cohort_flags(self, day, ncoh):
        """Boundary flags of the ``back_solve`` steps of ``day`` in the reference's call order
        (CalcSol.py:97-105: latest earlier release day first), population model with r_dur > 1."""
        out = []
        for j in range(ncoh - 1, -1, -1):
            sm = _abi.StepMeta()
            _lib.check(_lib.lib().pkb_result_cohort_meta(self.h, day, j, C.byref(sm)))
            out.append(bool(sm.flag))
        return out

    def spectral_steps(self):
        """Days whose chain step started from the stored spectrum of the state (option ``spectral``)."""
        return [d for d in range(self.ndays) if self.day_meta(d)[1].spec]

    def row_windows(self):
        """(first row, one past the last row) the step of each day computed, None where all rows."""
        out = []
        for d in range(self.ndays):
            sm = self.day_meta(d)[1]
            out.append((sm.wr0, sm.wr1) if sm.wr1 > sm.wr0 else None)
        return out

    def regions(self):
        """Per day (rows computed, columns computed, measured extent of the cells >= 1e-15) as
        ((wr0, wr1), (wc0, wc1), (er0, er1, ec0, ec1))."""
        out = []
        for d in range(self.ndays):
            sm = self.day_meta(d)[1]
            out.append(((sm.wr0, sm.wr1), (sm.wc0, sm.wc1), (sm.er0, sm.er1, sm.ec0, sm.ec1)))
        return out

    def radii(self):
        return [self.day_meta(d)[0].rad for d in range(self.ndays)]

    def dense(self, day):
        out = np.empty((self.dom_len, self.dom_len))
        _lib.check(_lib.lib().pkb_result_dense(self.h, day, _lib.dptr(out)))
        return out

    def pre(self, day):
        """Un-thresholded domain grid of one day (``solve(..., keep_pre=True)``): what ``ifft2`` returns before
        ``r_small_vals`` (CalcSol.py:189-190), or the cohort sum of CalcSol.py:322 for the population model."""
        out = np.empty((self.dom_len, self.dom_len))
        _lib.check(_lib.lib().pkb_result_pre(self.h, day, _lib.dptr(out)))
        return out

    def coo_arrays(self):
        """(day_offsets[ndays+1], rows, cols, vals) views on the result's pinned host buffers."""
        off, rows, cols, vals = _abi.c_ll_p(), _abi.c_int_p(), _abi.c_int_p(), _abi.c_double_p()
        _lib.check(_lib.lib().pkb_result_coo(self.h, C.byref(off), C.byref(rows), C.byref(cols), C.byref(vals)))
        o = np.ctypeslib.as_array(off, shape=(self.ndays + 1,))
        n = int(o[-1])
        if n == 0:
            return o, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0)
        return (o, np.ctypeslib.as_array(rows, shape=(n,)), np.ctypeslib.as_array(cols, shape=(n,)),
                np.ctypeslib.as_array(vals, shape=(n,)))

    def csr_arrays(self):
        """(day_offsets[ndays+1], row_offsets[ndays, dom_len], cols, vals) views on the result's pinned host buffers
        (``solve(..., want_coo='csr')``): the non-zeros of day d are ``[day_offsets[d], day_offsets[d+1])``, row r of that day
        starts at ``row_offsets[d, r]`` within them."""
        off, roff, cols, vals = _abi.c_ll_p(), _abi.c_ll_p(), _abi.c_int_p(), _abi.c_double_p()
        _lib.check(_lib.lib().pkb_result_csr(self.h, C.byref(off), C.byref(roff), C.byref(cols), C.byref(vals)))
        o = np.ctypeslib.as_array(off, shape=(self.ndays + 1,))
        ro = np.ctypeslib.as_array(roff, shape=(self.ndays, self.dom_len))
        n = int(o[-1])
        if n == 0:
            return o, ro, np.zeros(0, np.int32), np.zeros(0)
        return o, ro, np.ctypeslib.as_array(cols, shape=(n,)), np.ctypeslib.as_array(vals, shape=(n,))

    def csr_list(self):
        """One ``csr_matrix`` per day (what ``Run.main`` saves, Run.py:490-510)."""
        o, ro, cols, vals = self.csr_arrays()
        shp = (self.dom_len, self.dom_len)
        out = []
        for d in range(self.ndays):
            indptr = np.append(ro[d], o[d + 1] - o[d]).astype(np.int32)
            out.append(sparse.csr_matrix((vals[o[d]:o[d + 1]].copy(), cols[o[d]:o[d + 1]].copy(), indptr), shape=shp))
        return out

    def coo_list(self):
        """One ``coo_matrix`` per day (copies out of the pinned buffers)."""
        o, rows, cols, vals = self.coo_arrays()
        shp = (self.dom_len, self.dom_len)
        return [sparse.coo_matrix((vals[o[d]:o[d + 1]].copy(), (rows[o[d]:o[d + 1]].copy(), cols[o[d]:o[d + 1]].copy())),
                                  shape=shp) for d in range(self.ndays)]

    def sample(self, cells):
        cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 2)
        out = np.empty((self.ndays, cells.shape[0]))
        _lib.check(_lib.lib().pkb_result_sample(self.h, _lib.iptr(cells), cells.shape[0], _lib.dptr(out)))
        return out

    def device_ptr(self):
        p = C.c_void_p()
        _lib.check(_lib.lib().pkb_result_device_ptr(self.h, C.byref(p)))
        return p.value

    def close(self):
        if self.h:
            _lib.lib().pkb_result_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _solve_args(wind, ndays, hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res, prob_model, r_dur, r_number,
                r_dist, r_start, want_coo, want_dense, keep_device, wind_device_ptr, wind_shape, keep_pre=False,
                sprd_factor=None, sprd_drift=(-25., 15.)):
    """Fill a ``pkb_solve_args``; returns it with the arrays it points into (keep them alive)."""
    a = _abi.SolveArgs()
    if wind_device_ptr is not None:
        nd_wind, periods = int(wind_shape[0]), int(wind_shape[1])
        a.wind = C.c_void_p(int(wind_device_ptr))
        a.wind_on_device = 1
        keep = None
    else:
        keep = _lib.as_f64(wind)
        nd_wind, periods = keep.shape[0], keep.shape[1]
        a.wind = C.c_void_p(keep.ctypes.data)
        a.wind_on_device = 0
    a.nd_wind, a.periods, a.ndays = nd_wind, periods, int(ndays)
    a.day = PM._day_args(hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res, None, 0, False)
    a.prob_model = 1 if prob_model else 0
    a.r_dur = int(r_dur)
    a.r_number = float(r_number)
    w = _lib.as_f64(r_dist if r_dist is not None else [1.0 / max(int(r_dur), 1)] * max(int(r_dur), 1))
    a.r_dist = _lib.dptr(w)
    a.r_start = -1.0 if r_start is None else float(r_start)
    a.negval = 1e-8
    a.want_dense_host = 1 if want_dense else 0
    a.want_coo = 2 if want_coo == 'csr' else (1 if want_coo else 0)
    a.keep_dense_device = 1 if keep_device else 0
    a.keep_pre_device = 1 if keep_pre else 0
    if sprd_factor is not None:          # leading local-spread day (Bayes_Run.py:245-270, Bayes_MAP.py:247-277)
        a.sprd = 1
        a.sprd_factor = float(sprd_factor)
        a.sprd_drift[0], a.sprd_drift[1] = float(sprd_drift[0]), float(sprd_drift[1])
    return a, (keep, w)


def solve(wind, ndays, hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res, prob_model=True,
          r_dur=1, r_number=1.0, r_dist=None, r_start=None, want_coo=True, want_dense=False, keep_device=False,
          wind_device_ptr=None, wind_shape=None, device=None, keep_pre=False, sprd_factor=None, sprd_drift=(-25., 15.)):
    """Fused forward solve.  ``wind``: ndarray (nd_wind, periods, 3) of
    consecutive days (or None with ``wind_device_ptr``/``wind_shape`` for a
    wind array already resident on the device).  Returns a ``SolveResult``.

    ``sprd_factor`` (not None): the Bayes drivers' extra first day -- the release spreads locally for one day before
    the wind record starts (Bayes_Run.py:245-270, Bayes_MAP.py:247-277): a kernel mixed from two
    ``get_mvn_cdf_values`` blobs is prepended, the chain runs over ndays + 1 days and the first solution is dropped
    (Bayes_Run.py:288-296)."""
    a, keep = _solve_args(wind, ndays, hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res, prob_model, r_dur,
                          r_number, r_dist, r_start, want_coo, want_dense, keep_device, wind_device_ptr, wind_shape, keep_pre,
                          sprd_factor, sprd_drift)
    h = C.c_void_p()
    _lib.check(_lib.lib().pkb_solve(_lib.ctx(device).h, C.byref(a), C.byref(h)))
    del keep
    res = SolveResult(h)
    for d in range(res.ndays):
        PM._raise_for_status(res.day_meta(d)[0], d, (hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res))
    return res


def stack_wind(wind_data, days):
    """(nd, periods, 3) array of consecutive days for the fused solve."""
    if list(days) != list(range(days[0], days[0] + len(days))):
        raise ValueError('wind days must be consecutive integers (ParasitoidModel.py:178,214)')
    return np.stack([np.asarray(wind_data[d], dtype=float) for d in days])


def main(params):
    """Run one simulation (Run.py:388-520).  Returns ``modelsol``: a list with
    one sparse matrix per day (COO for the probability model, CSR for the
    population model), and writes ``params.outfile`` (.npz + .json) when
    ``params.OUTPUT``."""
    wind_data, days = PM.get_wind_data(*params.get_wind_params())
    ndays = min(params.ndays, len(days)) if params.ndays >= 0 else len(days)
    wind = stack_wind(wind_data, days)
    mp = params.get_model_params()
    tic = time.time()
    if params.PROB_MODEL:
        res = solve(wind, ndays, *mp, prob_model=True)
    else:
        dist = params.r_mthd()
        res = solve(wind, ndays, *mp, prob_model=False, r_dur=params.r_dur, r_number=params.r_number,
                    r_dist=[dist(d + 1) for d in range(params.r_dur)], r_start=params.r_start, want_coo='csr')
    modelsol = res.coo_list() if params.PROB_MODEL else res.csr_list()      # (get_populations returns CSR, CalcSol.py:323)
    res.close()
    print('Time elapsed: {0}'.format(time.time() - tic))

    if params.OUTPUT:
        out = {}
        for n, day in enumerate(days[:ndays]):                      # Run.py:490-510
            sol = modelsol[n].tocsr()
            out[str(day) + '_data'] = sol.data
            out[str(day) + '_ind'] = sol.indices
            out[str(day) + '_indptr'] = sol.indptr
        out['days'] = days[:ndays]
        dir_file = params.outfile.rsplit('/', 1)
        if len(dir_file) > 1 and not os.path.exists(dir_file[0]):
            os.makedirs(dir_file[0])
        np.savez(params.outfile, **out)
        with open(params.outfile + '.json', 'w') as fobj:
            param_dict = dict(params.__dict__)
            param_dict.pop('maps_key', None)
            json.dump(param_dict, fobj)
    return modelsol


if __name__ == '__main__':
    import sys
    _params = Params()
    if len(sys.argv[1:]) > 0:
        _params.cmd_line_chg(sys.argv[1:])
    main(_params)
