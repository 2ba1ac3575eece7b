"""Loader for libpkb200.so and small marshalling helpers.

There is exactly one backend: the in-tree CUDA library built by
``__graft_entry__.build()`` (nvcc, sm_100a).  If it is missing, or no CUDA
device is present, every entry point raises -- there is no CPU fallback and
nothing here imports ``oracle/``.
"""
import ctypes as C
import os
import threading

import numpy as np

from . import _abi

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libpkb200.so')

_LIB = None
_CTX = {}
_LOCK = threading.Lock()


class PkbError(RuntimeError):
    """Error reported by libpkb200 (code in ``.code``)."""

    def __init__(self, code, msg):
        RuntimeError.__init__(self, 'libpkb200 error {}: {}'.format(code, msg))
        self.code = code
        self.msg = msg


def lib():
    """The loaded library (loads it on first use)."""
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                'parasitoids_b200: {} not found. Build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a). This package has no CPU fallback.'.format(LIB_PATH))
        _LIB = _abi.declare(C.CDLL(LIB_PATH))
    return _LIB


def check(rc):
    if rc != 0:
        msg = lib().pkb_last_error()
        msg = msg.decode('utf-8', 'replace') if msg else ''
        if rc == _abi.PKB_EINVAL and ('must be positive' in msg or 'correlation must be' in msg):
            # Dmat's assertions (ParasitoidModel.py:276-278)
            raise AssertionError(msg)
        raise PkbError(rc, msg)


class Context(object):
    """One device context (stream, plan cache, buffer pools)."""

    def __init__(self, device=0):
        self.device = int(device)
        h = C.c_void_p()
        check(lib().pkb_create(self.device, C.byref(h)))
        self.h = h

    def set_option(self, key, value):
        check(lib().pkb_set_option(self.h, key.encode(), float(value)))

    def sync(self):
        check(lib().pkb_sync(self.h))

    def launch_count(self):
        return int(lib().pkb_launch_count(self.h))

    def timing(self):
        out = (C.c_double * 4)()
        check(lib().pkb_timing(self.h, out))
        return dict(phase1_ms=out[0], chain_ms=out[1], output_ms=out[2], total_ms=out[3])

    def mark(self, slot):
        check(lib().pkb_mark(self.h, int(slot)))

    def elapsed_ms(self, a, b):
        ms = C.c_double()
        check(lib().pkb_elapsed_ms(self.h, int(a), int(b), C.byref(ms)))
        return ms.value

    def profile(self, on):
        check(lib().pkb_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        check(lib().pkb_profile_reset(self.h))

    def profile_get(self, kernel):
        """(launch count, total device ms) of one kernel since the last reset."""
        n, ms = C.c_longlong(), C.c_double()
        check(lib().pkb_profile_get(self.h, kernel.encode(), C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def close(self):
        if self.h:
            lib().pkb_destroy(self.h)
            self.h = None


def default_device():
    """LOCAL_RANK under torchrun, else PKB_DEVICE, else 0."""
    for key in ('PKB_DEVICE', 'LOCAL_RANK'):
        if key in os.environ:
            return int(os.environ[key])
    return 0


def ctx(device=None):
    """Process-wide context for ``device`` (created on first use)."""
    dev = default_device() if device is None else int(device)
    with _LOCK:
        if dev not in _CTX:
            _CTX[dev] = Context(dev)
        return _CTX[dev]


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def dptr(a):
    return a.ctypes.data_as(_abi.c_double_p)


def iptr(a):
    return a.ctypes.data_as(_abi.c_int_p)
