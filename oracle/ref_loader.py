"""Load the UNMODIFIED reference modules (build container only).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

``/root/reference`` exists only in the build container, never on the GPU box,
so nothing reachable from ``pytest -m gpu``, ``smoke()`` or ``bench.py`` may
call this.  It is used by ``oracle/make_golden.py`` (to produce the committed
fixtures under tests/golden/) and by the container-only cross-checks in
tests/test_oracle.py (skipped when the tree is absent).

The reference's ``ParasitoidModel.py:22,340`` needs ``scipy.stats.mvn.mvnun``
(Fortran MVNDST), which current SciPy no longer ships.  SciPy does ship the
same Genz BVU algorithm as ``scipy.stats._qmvnt._bvn``; the shim below is the
three-line adapter described in SURVEY.md section 8c.
"""
import contextlib
import importlib
import io
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference in the build container; on the GPU box the copy that __graft_entry__.build() leaves under
# baseline/_ref/ (git-ignored, travels with the snapshot) -- used there ONLY by bench.py's reference arm
_SHIPPED = os.path.join(os.path.dirname(_HERE), 'baseline', '_ref')
REF_ROOT = os.environ.get('PARASITOIDS_REFERENCE') or ('/root/reference' if os.path.isfile('/root/reference/ParasitoidModel.py') else _SHIPPED)


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'ParasitoidModel.py'))


def _install_mvn_shim():
    import numpy as np
    import scipy.stats
    from scipy.stats import _qmvnt

    def mvnun(lower, upper, means, covar):
        lo = np.asarray(lower, dtype=float) - np.asarray(means, dtype=float)
        up = np.asarray(upper, dtype=float) - np.asarray(means, dtype=float)
        return _qmvnt._bvn(lo, up, np.asarray(covar, dtype=float)), 0

    shim = types.ModuleType('scipy.stats.mvn')
    shim.mvnun = mvnun
    sys.modules['scipy.stats.mvn'] = shim
    scipy.stats.mvn = shim


_BACK_SOLVE_BUG = 'bcksol_hat = fft2(sol,pad_shape)'
# fft2(A, filt) pads to A.shape + filt//2 (CalcSol.py:20-21); the intended
# re-FFT keeps pad_shape (cuda_lib.py:208-214), i.e. filt//2 == pad - dom.
_BACK_SOLVE_FIX = ('bcksol_hat = fft2(sol,2*(np.array(pad_shape)'
                   '-np.array(sol.shape)))')


def load(fixed_back_solve=False):
    """Return (ParasitoidModel, CalcSol, globalvars) reference modules.

    fixed_back_solve=True applies the documented one-line correction of the
    ``back_solve`` re-FFT shape defect (CalcSol.py:105, SURVEY.md section 8c
    defect 1) to the module source before executing it; everything else is
    the reference's code verbatim, executed from where it lies.
    """
    if not available():
        raise RuntimeError('reference tree not found at ' + REF_ROOT)
    _install_mvn_shim()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name in ('globalvars', 'CalcSol', 'ParasitoidModel'):
        sys.modules.pop(name, None)
    gv = importlib.import_module('globalvars')
    gv.cuda = False                              # CPU path is the oracle
    if fixed_back_solve:
        path = os.path.join(REF_ROOT, 'CalcSol.py')
        with open(path) as fobj:
            src = fobj.read()
        assert src.count(_BACK_SOLVE_BUG) == 1
        src = src.replace(_BACK_SOLVE_BUG, _BACK_SOLVE_FIX)
        cs = types.ModuleType('CalcSol')
        cs.__file__ = path
        sys.modules['CalcSol'] = cs
        exec(compile(src, path, 'exec'), cs.__dict__)
    else:
        cs = importlib.import_module('CalcSol')
    pm = importlib.import_module('ParasitoidModel')
    return pm, cs, gv


@contextlib.contextmanager
def quiet():
    """The reference prints one line per period (ParasitoidModel.py:437)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
