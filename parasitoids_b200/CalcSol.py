"""Drop-in for the reference's ``CalcSol`` module: the daily convolution chain
(CalcSol.py:11-324), computed on the GPU through ``cuda_lib.CudaSolve``.

``get_solutions`` / ``get_populations`` keep the reference signatures, mutate /
return the same scipy.sparse objects and follow the reference's CPU ordering
of threshold and cohort sum.  The primitives ``fft2`` / ``fftconv2`` / ``ifft2``
/ ``back_solve`` keep their call pattern, but the "spectrum" they pass around
is an opaque device-resident state (``SolHat``) rather than a complex ndarray:
the chain state never leaves the GPU.

There is no CPU fallback here (contrast CalcSol.py:160-174): if the CUDA
library or device is missing these functions raise.
"""
import numpy as np
from scipy import sparse

from . import cuda_lib


class SolHat(object):
    """Opaque chain state returned by ``fft2`` (stands in for the padded
    complex array of CalcSol.py:24)."""

    def __init__(self, solver, dom_shape):
        self.solver = solver
        self.dom_shape = tuple(int(s) for s in dom_shape)

    @property
    def shape(self):
        return self.solver.pad_shape


def fft2(A, filt_shape):
    """State for the sparse/dense solution ``A`` zero-padded by
    ``filt_shape//2`` (CalcSol.py:11-24)."""
    return SolHat(cuda_lib.CudaSolve(A, filt_shape), A.shape)


def ifft2(A_hat, Ashape):
    """(un-thresholded solution as COO, boundary flag) (CalcSol.py:28-41)."""
    dense, flag = A_hat.solver.get_solution(Ashape, raw=True, truncate=False)
    return sparse.coo_matrix(dense), flag


def fftconv2(A_hat, B):
    """``A_hat`` <- ``A_hat`` convolved with the odd-shaped filter ``B``, in
    place (CalcSol.py:45-66)."""
    A_hat.solver.fftconv2(B)


def back_solve(prev_spread, cursol_hat, dom_shape):
    """Cohorts of the earlier release days, un-thresholded, in emergence order
    (CalcSol.py:72-109, with the re-FFT kept at the padded shape -- see
    DESIGN.md "reference defects")."""
    return [sparse.coo_matrix(a) for a in cursol_hat.solver.back_solve_dense(prev_spread, dom_shape)]


def r_small_vals(A, prob_model=False, negval=1e-8):
    """Drop entries below ``negval``; with ``prob_model`` spread the missing
    mass uniformly over the survivors (CalcSol.py:112-136).  Host-side helper
    for callers that hold a scipy matrix; the chain applies the same rule on
    the device (k_emit_dense / k_emit_population)."""
    if not sparse.isspmatrix_coo(A):
        A = sparse.coo_matrix(A)
    keep = ~(A.data < negval)
    out = sparse.coo_matrix((A.data[keep], (A.row[keep], A.col[keep])), A.shape)
    if prob_model:
        out.data += (1 - out.data.sum()) / out.data.size
    return out


def get_solutions(modelsol, pmf_list, days, ndays, dom_len, max_shape, details=None):
    """Append the solutions of days 2..ndays to ``modelsol`` (CalcSol.py:140-201).

    ``details`` (optional dict, not in the reference): 'flags' per step and,
    if details.get('want_pre'), the dense un-thresholded solutions in 'pre'."""
    D = [dom_len, dom_len]
    solver = cuda_lib.CudaSolve(modelsol[0], max_shape)
    flags, pre = [], []
    want_pre = bool(details and details.get('want_pre'))
    try:
        for n, _day in enumerate(days[1:ndays]):
            solver.fftconv2(pmf_list[n + 1].tocsr(), False)
            if want_pre:
                raw, _ = solver.get_solution(D, raw=True, truncate=False)
                pre.append(raw)
            sol, flag = solver.get_solution(D, prob_model=True, truncate=True)
            modelsol.append(sparse.coo_matrix(sol))
            flags.append(flag)
    finally:
        solver.close()
    if details is not None:
        details['flags'] = flags
        details['pre'] = pre


def get_populations(r_spread, pmf_list, days, ndays, dom_len, max_shape, r_dur, r_number, dist, details=None):
    """Expected wasp numbers per day as a list of CSR matrices (CalcSol.py:205-324)."""
    D = [dom_len, dom_len]
    popmodel = []
    flags, pre = [], []
    want_pre = bool(details and details.get('want_pre'))
    w = [float(dist(d + 1)) for d in range(r_dur)]

    def emit(solver, ncoh, extra, add_centre, first=False):
        res = solver.population(w[:ncoh], r_number, extra, add_centre, first_day=first, want_pre=want_pre)
        if want_pre:
            pre.append(res[1])
            res = res[0]
        popmodel.append(sparse.csr_matrix(res))

    solver = cuda_lib.CudaSolve(r_spread[0], max_shape)
    try:
        emit(solver, 1, r_number * (1 - w[0]), True, first=True)
        for day in range(1, r_dur):
            solver.close()
            solver = cuda_lib.CudaSolve(r_spread[day], max_shape)
            solver.back_solve_dense(r_spread[:day], D, fetch=False)
            emit(solver, day + 1, (1 - sum(w[:day + 1])) * r_number, True)
        for n, _day in enumerate(days[r_dur:ndays]):
            solver.fftconv2(pmf_list[n + r_dur].tocsr(), False)
            _, flag = solver.get_solution(D, raw=True, truncate=True)
            flags.append(flag)
            solver.back_solve_dense(r_spread[:-1], D, fetch=False)
            emit(solver, r_dur, 0.0, False)
    finally:
        solver.close()
    if details is not None:
        details['flags'] = flags
        details['pre'] = pre
    return popmodel
