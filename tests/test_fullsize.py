"""BASELINE.json's full-size configuration (C4: 4097 x 4097 domain, 60 flight
days) on the GPU: size-independent properties of the whole solve, and a direct
comparison of the first chain steps against the oracle at full size (the
oracle needs ~3 s per 4279^2 step, so only two steps are compared densely)."""
import os
import sys

import numpy as np
import pytest
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import helpers as H        # noqa: E402


def _c4():
    import bench
    return bench.load_workload('synthetic_4097x4097_60d')


@pytest.mark.gpu
def test_c4_whole_solve_properties(gpu):
    """60 days at 4097^2: every day's probability grid sums to 1 within 1e-12, is
    non-negative, thresholded at 1e-8; no boundary flag trips (the
    C4 wind keeps the mass inside the domain); kernel radii as the survey measured."""
    wind, wind_data, days, rad_dist, rad_res = _c4()
    res = gpu.Run.solve(wind, len(days), H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 30, rad_dist, rad_res,
                        prob_model=True, want_coo=True)
    try:
        assert res.dom_len == 4097 and res.P == 4097 + res.max_shape // 2
        assert not any(res.flags())
        radii = res.radii()
        assert 60 <= min(radii) and max(radii) <= 190
        off, rows, cols, vals = res.coo_arrays()
        assert vals.min() >= 1e-8 * (1 - 1e-12)
        nnz = np.diff(off)
        assert np.all(nnz[1:] > nnz[0]), 'support must grow from the first day on'
        for d in range(len(days)):
            v = vals[off[d]:off[d + 1]]
            assert abs(np.sum(v, dtype=np.longdouble) - 1) < H.MASS, 'day %d mass' % d
        # centre of mass drifts with the mean wind (+x, -y => columns grow, rows grow: y is flipped)
        last = slice(off[-2], off[-1])
        cm_col = float((cols[last] * vals[last]).sum())
        cm_row = float((rows[last] * vals[last]).sum())
        assert cm_col > rad_res + 50 and cm_row > rad_res + 20
    finally:
        res.close()


@pytest.mark.gpu
def test_c4_first_steps_against_oracle(gpu):
    """Days 1-3 of C4 at full size: kernels from the GPU, chain steps from the GPU
    and from the oracle (pocketfft on the 4279^2 reference torus), dense
    pre-threshold grids compared at the north-star tolerances."""
    from oracle import cs_oracle as CO
    wind, wind_data, days, rad_dist, rad_res = _c4()
    nd = 3
    args = [(d, wind_data, H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 30, rad_dist, rad_res) for d in days[:nd]]
    pmfs = gpu.PM.prob_mass_batch(args)
    D = 2 * rad_res + 1
    ms = [max(p.shape[0] for p in pmfs)] * 2
    sol_ref = [H.recentre(pmfs[0], rad_res)]
    det = {}
    CO.get_solutions(sol_ref, pmfs, days, nd, D, ms, details=det)
    sol_gpu = [H.recentre(pmfs[0], rad_res)]
    dg = {'want_pre': True}
    gpu.CS.get_solutions(sol_gpu, pmfs, days, nd, D, ms, details=dg)
    assert dg['flags'] == det['flags'] == [False, False]
    for n in range(nd - 1):
        H.assert_parity(dg['pre'][n], det['pre'][n], what='C4 day %d pre-threshold' % (n + 2))
        got, ref = sol_gpu[n + 1].toarray(), sol_ref[n + 1].toarray()
        H.assert_thresholded_parity(got, ref, what='C4 day %d' % (n + 2))
        assert abs(got.sum() - 1) < H.MASS


@pytest.mark.gpu
def test_c4_convolution_commutes(gpu):
    """Un-flagged chain steps are circular convolutions mod P, so they commute:
    (A * K1) * K2 == (A * K2) * K1 at full size, to rounding."""
    rng = np.random.default_rng(5)
    D, k = 4097, 361
    A = np.zeros((D, D))
    A[1800:2300, 1700:2400] = rng.random((500, 700))
    A /= A.sum()
    K1 = rng.random((k, k)); K1 /= K1.sum()
    K2 = np.zeros((k, k)); K2[100:260, 60:300] = rng.random((160, 240)); K2 /= K2.sum()
    outs = []
    for order in ((K1, K2), (K2, K1)):
        s = gpu.cuda_lib.CudaSolve(sparse.coo_matrix(A), [k, k])
        try:
            for K in order:
                s.fftconv2(sparse.csr_matrix(K))
            dense, flag = s.get_solution([D, D], raw=True, truncate=False)
            assert not flag
            outs.append(dense)
        finally:
            s.close()
    assert abs(outs[0].sum() - 1) < 1e-12
    assert np.abs(outs[0] - outs[1]).max() < 1e-18 + 1e-12 * outs[0].max()
