"""CPU oracle, phase 2: FFT convolution chain (CalcSol.py restated).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  numpy + scipy.fft
(pocketfft, complex128), the same transform library the reference reaches
through scipy.fftpack.

Deviations from the reference, all documented in DESIGN.md:
  * ``back_solve`` re-FFTs at the SAME padded shape after a boundary flag
    (CalcSol.py:105 pads to D + P//2 instead of P and then fails to broadcast;
    the intended semantics are those of cuda_lib.py:208-214).  Pass
    ``fixed=False`` to reproduce the reference defect for small inputs.
  * ``np.sum(generator)`` (CalcSol.py:271,303,322) is restated with the
    builtin ``sum`` (same left-to-right order).
Every routine can also hand back the dense pre-threshold grids the parity
tests compare on.
"""
import numpy as np
from scipy import sparse
from scipy import fft as sfft

from .pm_oracle import r_small_vals  # noqa: F401  (CalcSol.py:112-136)

# Threads pocketfft may use for the 2-D transforms (scipy.fft ``workers``).  None = 1 = what scipy.fftpack does in
# the reference; the full-size parity tests set -1: every 1-D transform is computed by the same code whichever
# thread runs it, so the results are bit-identical and only the wall time changes.
WORKERS = None


def pad_shape_of(shape, filt_shape):
    """CalcSol.py:20-21."""
    return tuple(int(s) + int(f) // 2 for s, f in zip(shape, filt_shape))


def fft2(A, filt_shape):
    """CalcSol.py:11-24 -- zero-pad to A.shape + filt_shape//2, complex FFT."""
    A = sparse.coo_matrix(A)
    P = pad_shape_of(A.shape, filt_shape)
    buf = np.zeros(P)
    buf[:A.shape[0], :A.shape[1]] = A.toarray()
    return sfft.fft2(buf, workers=WORKERS)


def ifft2_dense(A_hat, Ashape):
    """CalcSol.py:28-41 without the COO conversion.

    Returns (dense real P x P, flag)."""
    A = sfft.ifft2(A_hat, workers=WORKERS).real
    r, c = int(Ashape[0]), int(Ashape[1])
    pads = []
    for blk in (A[r:, c:], A[:r, c:], A[r:, :c]):
        if blk.size:
            pads.append(blk.max())
    flag = bool(pads) and max(pads) > 1e-8
    return A, flag


def ifft2(A_hat, Ashape):
    """CalcSol.py:28-41."""
    A, flag = ifft2_dense(A_hat, Ashape)
    return sparse.coo_matrix(A[:int(Ashape[0]), :int(Ashape[1])]), flag


def wrap_kernel(B, pad_shape):
    """CalcSol.py:58-64 -- centre of the odd-shaped kernel to [0, 0] with
    wrap-around on the padded torus."""
    B = np.asarray(B.toarray() if sparse.issparse(B) else B, dtype=float)
    m0, m1 = B.shape[0] // 2, B.shape[1] // 2
    if m0 == 0 or m1 == 0:
        raise ValueError('1x1 kernels are not supported by fftconv2 '
                         '(negative-zero slices, CalcSol.py:62-64)')
    out = np.zeros(pad_shape)
    out[:m0 + 1, :m1 + 1] = B[m0:, m1:]
    out[:m0 + 1, -m1:] = B[m0:, :m1]
    out[-m0:, -m1:] = B[:m0, :m1]
    out[-m0:, :m1 + 1] = B[:m0, m1:]
    return out


def fftconv2(A_hat, B):
    """CalcSol.py:45-66 -- A_hat *= FFT(wrap-shifted B), in place."""
    A_hat *= sfft.fft2(wrap_kernel(B, A_hat.shape), workers=WORKERS)


def back_solve(prev_spread, cursol_hat, dom_shape, fixed=True, dense=False):
    """CalcSol.py:72-109 -- cohorts of earlier release days."""
    out = []
    hat = np.array(cursol_hat)
    P = cursol_hat.shape
    for B in prev_spread[::-1]:
        hat = sfft.fft2(wrap_kernel(B, P), workers=WORKERS) * hat
        A, flag = ifft2_dense(hat, dom_shape)
        sol = A[:int(dom_shape[0]), :int(dom_shape[1])]
        if flag:
            if fixed:
                buf = np.zeros(P)
                buf[:sol.shape[0], :sol.shape[1]] = sol
                hat = sfft.fft2(buf, workers=WORKERS)
            else:
                hat = fft2(sparse.coo_matrix(sol), P)      # reference defect
        out.append(sol.copy() if dense else sparse.coo_matrix(sol))
    return out[::-1]


def get_solutions(modelsol, pmf_list, days, ndays, dom_len, max_shape,
                  details=None):
    """CalcSol.py:140-201 (CPU branch).  Appends to modelsol in place.

    details (optional dict): 'pre' -> list of dense D x D un-thresholded
    solutions for days 2..ndays, 'flags' -> list of bool."""
    D = [dom_len, dom_len]
    hat = fft2(modelsol[0], max_shape)
    pre, flags = [], []
    for n, _day in enumerate(days[1:ndays]):
        fftconv2(hat, pmf_list[n + 1].tocsr())
        A, flag = ifft2_dense(hat, D)
        dom = A[:dom_len, :dom_len]
        modelsol.append(r_small_vals(sparse.coo_matrix(dom), prob_model=True))
        pre.append(dom.copy())
        flags.append(flag)
        if flag:
            buf = np.zeros(hat.shape)
            buf[:dom_len, :dom_len] = dom
            hat = sfft.fft2(buf, workers=WORKERS)
    if details is not None:
        details['pre'] = pre
        details['flags'] = flags


def get_populations(r_spread, pmf_list, days, ndays, dom_len, max_shape,
                    r_dur, r_number, dist, fixed=True, details=None):
    """CalcSol.py:205-324 (CPU branch).  Returns list of CSR, one per day.

    details (optional dict): 'pre' -> dense D x D weighted cohort sums before
    r_small_vals (without the centre remainder), 'flags' -> per post-release
    day flag of the leading cohort."""
    D = [dom_len, dom_len]
    mid = dom_len // 2
    cur = [0 for _ in range(r_dur)]
    pop = []
    pre, flags = [], []

    first = r_small_vals(r_spread[0]).tocsr() * r_number * dist(1)
    first = first.tolil()
    first[mid, mid] += r_number * (1 - dist(1))
    pop.append(first.tocsr())
    pre.append(np.asarray(sparse.coo_matrix(r_spread[0]).toarray()) * r_number * dist(1))
    cur[0] = np.asarray(sparse.coo_matrix(r_spread[0]).toarray())

    if r_dur == 1:
        hat = fft2(r_spread[0], max_shape)
    for day in range(1, r_dur):
        hat = fft2(r_spread[day], max_shape)
        cur[day] = np.asarray(sparse.coo_matrix(r_spread[day]).toarray())
        cur[:day] = back_solve(r_spread[:day], hat, D, fixed=fixed, dense=True)
        tot = sum(cur[d] * dist(d + 1) for d in range(day + 1)) * r_number
        pre.append(np.array(tot))
        sol = r_small_vals(sparse.coo_matrix(tot)).tolil()
        sol[mid, mid] += (1 - sum(dist(d + 1) for d in range(day + 1))) * r_number
        pop.append(sol.tocsr())
    for n, _day in enumerate(days[r_dur:ndays]):
        fftconv2(hat, pmf_list[n + r_dur].tocsr())
        A, flag = ifft2_dense(hat, D)
        cur[-1] = A[:dom_len, :dom_len].copy()
        flags.append(flag)
        if flag:
            buf = np.zeros(hat.shape)
            buf[:dom_len, :dom_len] = cur[-1]
            hat = sfft.fft2(buf, workers=WORKERS)
        cur[:-1] = back_solve(r_spread[:-1], hat, D, fixed=fixed, dense=True)
        tot = sum(cur[d] * dist(d + 1) for d in range(r_dur)) * r_number
        pre.append(np.array(tot))
        pop.append(r_small_vals(sparse.coo_matrix(tot)).tocsr())
    if details is not None:
        details['pre'] = pre
        details['flags'] = flags
    return pop
