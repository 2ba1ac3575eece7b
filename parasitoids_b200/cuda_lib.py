"""Drop-in for the reference's ``cuda_lib`` module: the GPU backend seam that
``CalcSol.get_solutions`` / ``get_populations`` select when ``globalvars.cuda``
is set (CalcSol.py:160-186, 241-288).

``CudaSolve`` keeps the reference's constructor and its three methods
(cuda_lib.py:16-221) but computes in fp64 through libpkb200's sm_100a kernels
instead of Reikna/PyCUDA complex64, so its results match the reference's CPU
path to rounding.  Differences a caller can observe: values are float64, and
domains / filters must be square (every reference caller's are).
"""
import ctypes as C

import numpy as np
from scipy import sparse

from . import _abi
from . import _lib


def _dense(A):
    if sparse.issparse(A):
        A = A.toarray()
    return _lib.as_f64(A)


def _square_side(shape, what):
    if len(shape) != 2 or int(shape[0]) != int(shape[1]):
        raise ValueError('{} must be square, got shape {}'.format(what, tuple(shape)))
    return int(shape[0])


class CudaSolve(object):

    def __init__(self, A, max_shape):
        """Initialise the solver with the solution after the first day.

        Args:
            A: first day's spread, sparse (or dense) square matrix
            max_shape: shape of the largest filter; the padded torus is
                A.shape + max_shape//2 (cuda_lib.py:25-28)"""
        D = _square_side(A.shape, 'A')
        ms = np.array(max_shape).astype(int).ravel()
        if ms.size == 1:
            ms = np.array([ms[0], ms[0]])
        if ms[0] != ms[1]:
            raise ValueError('max_shape must be square, got {}'.format(tuple(ms)))
        self.dom_len = D
        self.pad_shape = (D + int(ms[0]) // 2, D + int(ms[1]) // 2)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().pkb_chain_create(_lib.ctx().h, D, int(ms[0]), C.byref(self._h)))
        _lib.check(_lib.lib().pkb_chain_set_state(self._h, _lib.dptr(_dense(A))))

    # -- reference methods ----------------------------------------------------
    def fftconv2(self, B, mem_print=False):
        """Update the current solution with filter B (cuda_lib.py:58-94).
        B: square, odd-sided, centre of the filter at B[k//2, k//2]."""
        k = _square_side(B.shape, 'B')
        if k < 3:
            raise ValueError('filters must be at least 3x3 (CalcSol.py:62-64 breaks on 1x1 filters)')
        _lib.check(_lib.lib().pkb_chain_conv(self._h, _lib.dptr(_dense(B)), k))

    def get_cursol(self, dom_shape, negval=1e-8):
        """Current solution with entries below ``negval`` removed; truncates
        the state to the domain if anything above ``negval`` has reached the
        padding (cuda_lib.py:98-140).  Returns a COO matrix."""
        self._check_dom(dom_shape)
        out = np.empty((self.dom_len, self.dom_len))
        meta = _abi.StepMeta()
        _lib.check(_lib.lib().pkb_chain_get_cursol(self._h, float(negval), 1, 1, _lib.dptr(out), C.byref(meta)))
        self.last_flag = bool(meta.flag)
        return sparse.coo_matrix(out)

    def back_solve(self, prev_spread, dom_shape, negval=1e-8):
        """Convolve the current solution progressively with the filters of
        ``prev_spread`` in reverse order (cuda_lib.py:145-221).  Returns COO
        matrices in order of emergence, each thresholded at ``negval``."""
        return [sparse.coo_matrix(a) for a in self._back_solve(prev_spread, dom_shape, float(negval))]

    # -- extensions used by CalcSol (CPU-path ordering of the threshold) ------
    def get_solution(self, dom_shape, negval=1e-8, prob_model=True, truncate=True, raw=False):
        """(dense solution, boundary flag).  raw: un-thresholded domain values
        (CalcSol.ifft2); else ``r_small_vals`` applied on the device."""
        self._check_dom(dom_shape)
        out = np.empty((self.dom_len, self.dom_len))
        meta = _abi.StepMeta()
        mode = 0 if raw else (2 if prob_model else 1)
        _lib.check(_lib.lib().pkb_chain_get_cursol(self._h, float(negval), mode, 1 if truncate else 0,
                                                   _lib.dptr(out), C.byref(meta)))
        self.last_flag = bool(meta.flag)
        return out, bool(meta.flag)

    def back_solve_dense(self, prev_spread, dom_shape, fetch=True):
        """Un-thresholded cohorts (CalcSol.back_solve ordering, CalcSol.py:72-109);
        fetch=False leaves them on the device for ``population``."""
        return self._back_solve(prev_spread, dom_shape, -1.0, fetch)

    def population(self, weights, r_number, centre_extra=0.0, add_centre=False, negval=1e-8, first_day=False,
                   want_pre=False):
        """Cohort superposition on the device (CalcSol.py:236-237,271-274,303-306,322-323)."""
        w = _lib.as_f64(weights)
        out = np.empty((self.dom_len, self.dom_len))
        pre = np.empty((self.dom_len, self.dom_len)) if want_pre else None
        _lib.check(_lib.lib().pkb_chain_population(
            self._h, w.size, _lib.dptr(w), float(r_number), float(centre_extra), 1 if add_centre else 0, float(negval),
            1 if first_day else 0, _lib.dptr(out), _lib.dptr(pre) if want_pre else None))
        return (out, pre) if want_pre else out

    def state(self):
        """Full padded state (diagnostics)."""
        P = self.pad_shape[0]
        out = np.empty((P, P))
        _lib.check(_lib.lib().pkb_chain_get_state(self._h, _lib.dptr(out)))
        return out

    def dims(self):
        D, P, N = C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib().pkb_chain_dims(self._h, C.byref(D), C.byref(P), C.byref(N)))
        return D.value, P.value, N.value

    # -- internals --------------------------------------------------------------
    def _check_dom(self, dom_shape):
        if int(dom_shape[0]) != self.dom_len or int(dom_shape[1]) != self.dom_len:
            raise ValueError('dom_shape {} does not match the solver domain {}'.format(tuple(dom_shape), self.dom_len))

    def _back_solve(self, prev_spread, dom_shape, threshold, fetch=True):
        self._check_dom(dom_shape)
        nf = len(prev_spread)
        if nf == 0:
            return []
        mats = [_dense(B) for B in prev_spread]
        ks = np.array([_square_side(m.shape, 'filter') for m in mats], dtype=np.int32)
        ptrs = (_abi.c_double_p * nf)(*[_lib.dptr(m) for m in mats])
        out = np.empty((nf, self.dom_len, self.dom_len)) if fetch else None
        flags = np.zeros(nf, dtype=np.int32)
        _lib.check(_lib.lib().pkb_chain_back_solve(self._h, ptrs, _lib.iptr(ks), nf, threshold,
                                                   _lib.dptr(out) if fetch else None, _lib.iptr(flags)))
        self.last_back_flags = [bool(f) for f in flags]
        return [out[j] for j in range(nf)] if fetch else []

    def close(self):
        if getattr(self, '_h', None):
            _lib.lib().pkb_chain_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
