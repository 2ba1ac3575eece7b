"""Test configuration.

Two backends appear in the parity tests through the ``pkb`` fixture:

* ``gpu``  -- the product: parasitoids_b200/libpkb200.so on a real B200
  (marked ``gpu``; selected with ``-m gpu``).
* ``emul`` -- TEST INFRASTRUCTURE: the same kernel + host sources compiled
  with g++ against tests/emul/emul_cuda.h (a fiber emulation of CUDA blocks)
  so that the indexing / control-flow logic of every kernel and all of the
  Python host layer are exercised in the GPU-less build container.  It is
  injected into ``parasitoids_b200._lib`` only here; the package itself has no
  way to select it.
"""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMUL_DIR = os.path.join(ROOT, 'tests', 'emul')
EMUL_SO = os.path.join(EMUL_DIR, 'libpkb200_emul.so')
CSRC = os.path.join(ROOT, 'parasitoids_b200', 'csrc')
GOLD = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'slow: long-running CPU test')


def build_emul():
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(EMUL_DIR, f) for f in ('emul_cuda.h', 'emul_cuda.cpp')] + \
           [os.path.join(ROOT, 'include', 'pkb200.h')]
    if os.path.isfile(EMUL_SO) and all(os.path.getmtime(EMUL_SO) >= os.path.getmtime(s) for s in srcs):
        return EMUL_SO
    cmd = ['g++', '-O2', '-std=c++17', '-DPKB_EMUL', '-ffp-contract=off', '-I' + EMUL_DIR, '-I' + CSRC,
           '-x', 'c++', os.path.join(CSRC, 'pkb200.cu'), '-x', 'c++', os.path.join(EMUL_DIR, 'emul_cuda.cpp'),
           '-shared', '-fPIC', '-o', EMUL_SO, '-lpthread']
    subprocess.run(cmd, check=True, cwd=ROOT)
    return EMUL_SO


class Backend(object):
    def __init__(self, kind):
        self.kind = kind
        from parasitoids_b200 import ParasitoidModel, CalcSol, cuda_lib, Run, _lib, _abi
        self.PM, self.CS, self.cuda_lib, self.Run, self._lib, self._abi = ParasitoidModel, CalcSol, cuda_lib, Run, _lib, _abi

    @property
    def is_gpu(self):
        return self.kind == 'gpu'


def _activate(kind):
    from parasitoids_b200 import _lib, _abi
    _lib._CTX.clear()
    if kind == 'emul':
        _lib._LIB = _abi.declare(ctypes.CDLL(build_emul()))
    else:
        _lib._LIB = None
        _lib.lib()


@pytest.fixture(params=[pytest.param('emul'), pytest.param('gpu', marks=pytest.mark.gpu)])
def pkb(request):
    """Backend under test: emulated (CPU container) or the real library (-m gpu)."""
    _activate(request.param)
    yield Backend(request.param)
    from parasitoids_b200 import _lib
    _lib._CTX.clear()
    _lib._LIB = None


@pytest.fixture
def gpu():
    """The real library only (use together with @pytest.mark.gpu)."""
    _activate('gpu')
    yield Backend('gpu')
    from parasitoids_b200 import _lib
    _lib._CTX.clear()
    _lib._LIB = None


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLD, name + '.npz'))
    return load
