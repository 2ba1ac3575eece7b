"""Phase 2 (CalcSol / cuda_lib drop-in and the fused solve) against the
reference's known-answer tests, its golden vectors and the oracle.  Runs on
the emulated backend in the CPU container and on the real library with -m gpu."""
import warnings

import numpy as np
import pytest
from scipy import signal, sparse

from oracle import cs_oracle as CO
from oracle import pm_oracle as PO
import helpers as H


# ---- the reference's own tests (tests/test_CalcSol.py), restated -------------
def _two_arrays():
    A = np.outer(range(10), range(1, 11))
    B = np.outer(range(4, -1, -1), range(8, -1, -2))
    return A, B


def _many_arrays():
    out = []
    for data in (np.outer(range(5), np.arange(.1, .6, .1)), np.outer(np.arange(0, 2.5, 0.5), np.ones(5)),
                 np.outer(range(5, 0, -1), np.arange(.1, .6, .1)), np.outer(np.arange(1, 0, -.2), np.arange(0, 2.5, 0.5))):
        M = np.zeros((55, 55))
        M[25:30, 25:30] = data
        out.append(M)
    return out


def test_fftconv2(pkb):
    """tests/test_CalcSol.py:75-83."""
    CS = pkb.CS
    A, B = _two_arrays()
    A0, B0 = A.copy(), B.copy()
    A_hat = CS.fft2(sparse.coo_matrix(A), np.array(B.shape))
    before, _ = CS.ifft2(A_hat, A.shape)
    CS.fftconv2(A_hat, sparse.csr_matrix(B))
    after, _ = CS.ifft2(A_hat, A.shape)
    assert not np.all(before.toarray() == after.toarray())
    assert np.all(B == B0) and np.all(A == A0)
    assert np.allclose(after.toarray(), signal.convolve2d(A, B, 'same'), rtol=1e-12, atol=1e-10)


def test_convolve_same(pkb):
    """tests/test_CalcSol.py:85-98."""
    CS = pkb.CS
    A, B = _two_arrays()
    A_hat = CS.fft2(sparse.coo_matrix(A), np.array([A.shape[0] + 6, A.shape[1] + 6]))
    CS.fftconv2(A_hat, sparse.csr_matrix(B))
    C, flag = CS.ifft2(A_hat, A.shape)
    C = C.toarray()
    assert not np.iscomplexobj(C)
    assert np.allclose(C, signal.fftconvolve(A, B, 'same'), rtol=1e-12, atol=1e-10)


def test_cuda_convolve(pkb):
    """tests/test_CalcSol.py:100-113 (fp64, so without the float32 tolerance)."""
    A, B = _two_arrays()
    solver = pkb.cuda_lib.CudaSolve(sparse.coo_matrix(A), np.array(A.shape) + 6)
    solver.fftconv2(sparse.csr_matrix(B))
    C = solver.get_cursol(A.shape)
    assert sparse.isspmatrix_coo(C)
    assert np.allclose(C.toarray(), signal.fftconvolve(A, B, 'same'), rtol=1e-12, atol=1e-10)


def _abcd_reference():
    A, B, C, D = _many_arrays()
    B_hat = CO.fft2(sparse.coo_matrix(B), A.shape)
    CO.fftconv2(B_hat, sparse.csr_matrix(C))
    CO.fftconv2(B_hat, sparse.csr_matrix(D))
    BCD = CO.ifft2(B_hat, B.shape)[0].toarray()
    A_hat = CO.fft2(sparse.coo_matrix(A), A.shape)
    for M in (B, C, D):
        CO.fftconv2(A_hat, sparse.csr_matrix(M))
    ABCD = CO.ifft2(A_hat, A.shape)[0].toarray()
    return BCD, ABCD


def test_back_solve(pkb):
    """tests/test_CalcSol.py:115-139."""
    CS = pkb.CS
    A, B, C, D = _many_arrays()
    C_hat = CS.fft2(sparse.coo_matrix(C), A.shape)
    CS.fftconv2(C_hat, sparse.csr_matrix(D))
    bck = CS.back_solve([sparse.csr_matrix(A), sparse.csr_matrix(B)], C_hat, A.shape)
    BCD, ABCD = _abcd_reference()
    assert np.allclose(bck[1].toarray(), BCD, rtol=1e-12, atol=1e-10)
    assert np.allclose(bck[0].toarray(), ABCD, rtol=1e-12, atol=1e-10)


def test_cuda_back_solve(pkb):
    """tests/test_CalcSol.py:141-171 (fp64: tolerances 1e-4/1e-3 -> 1e-10)."""
    A, B, C, D = _many_arrays()
    solver = pkb.cuda_lib.CudaSolve(sparse.coo_matrix(C), A.shape)
    solver.fftconv2(sparse.csr_matrix(D))
    bck = solver.back_solve([sparse.csr_matrix(A), sparse.csr_matrix(B)], A.shape)
    BCD, ABCD = _abcd_reference()
    # cuda_lib thresholds each cohort at 1e-8 (cuda_lib.py:195-197)
    assert np.allclose(bck[1].toarray(), np.where(BCD > 1e-8, BCD, 0), rtol=1e-12, atol=1e-10)
    assert np.allclose(bck[0].toarray(), np.where(ABCD > 1e-8, ABCD, 0), rtol=1e-12, atol=1e-10)


def test_stencil_path_matches_fft_path(pkb):
    """Small-support filters go through the direct shared-memory stencil; both
    paths must agree with the oracle's circular convolution."""
    rng = np.random.default_rng(3)
    A = rng.random((40, 40))
    B = rng.random((5, 5))
    ref_hat = CO.fft2(sparse.coo_matrix(A), (21, 21))
    CO.fftconv2(ref_hat, sparse.csr_matrix(B))
    ref, _ = CO.ifft2_dense(ref_hat, A.shape)
    for radius in (-1, 3):
        pkb._lib.ctx().set_option('stencil_max_radius', radius)
        solver = pkb.cuda_lib.CudaSolve(sparse.coo_matrix(A), (21, 21))
        solver.fftconv2(sparse.csr_matrix(B))
        assert np.abs(solver.state() - ref).max() < 1e-12, radius
        solver.close()
    pkb._lib.ctx().set_option('stencil_max_radius', 3)


def test_get_cursol_uses_the_callers_negval(pkb):
    """cuda_lib.get_cursol(dom_shape, negval) keeps entries strictly above negval and raises the boundary flag -- and
    truncates -- when something above NEGVAL (not a fixed 1e-8) has reached the padding (cuda_lib.py:117-136)."""
    D = 21
    A = np.zeros((D, D))
    A[10, D - 1] = 1.0                                   # on the right edge: one third of the mass spills into the padding
    B = np.full((3, 3), 1.0 / 9.0)
    v = 1.0 / 9.0                                        # every touched cell holds exactly this (direct stencil, one term each)
    for negval, kept, flag in ((1e-8, 6, True), (v, 0, False), (0.2, 0, False)):
        s = pkb.cuda_lib.CudaSolve(sparse.coo_matrix(A), (3, 3))
        try:
            s.fftconv2(sparse.csr_matrix(B))
            sol = s.get_cursol([D, D], negval)
            assert sol.nnz == kept, (negval, sol.nnz)
            assert s.last_flag == flag, (negval, s.last_flag)
            if kept:
                assert np.all(sol.data > negval) and np.allclose(sol.data, v, rtol=0, atol=1e-16)
        finally:
            s.close()


def test_filter_limits(pkb):
    A = np.ones((9, 9))
    solver = pkb.cuda_lib.CudaSolve(sparse.coo_matrix(A), (5, 5))
    with pytest.raises(pkb._lib.PkbError):
        solver.fftconv2(sparse.csr_matrix(np.ones((9, 9))))          # support radius 4 > max_shape//2 = 2
    with pytest.raises(ValueError):
        solver.fftconv2(sparse.csr_matrix(np.ones((1, 1))))
    with pytest.raises(ValueError):
        pkb.cuda_lib.CudaSolve(sparse.coo_matrix(np.ones((4, 5))), (3, 3))


# ---- golden chains (reference get_solutions / get_populations) ---------------
def _small_pmfs():
    z = H.load('pm_small')
    days = [int(d) for d in z['days']]
    return z, days, [H.coo(z, 'd%d_pmf' % d) for d in days]


def test_get_solutions_small(pkb):
    z, days, pmfs = _small_pmfs()
    g = H.load('chain_small')
    n = 10
    rr, D, ms = int(z['rad_res']), int(g['dom_len']), g['max_shape']
    sol = [H.recentre(pmfs[0], rr)]
    det = {'want_pre': True}
    pkb.CS.get_solutions(sol, pmfs[:n], days[:n], n, D, ms, details=det)
    assert len(sol) == n
    assert det['flags'] == [bool(f) for f in g['prob_flags']]
    for i in range(1, n):
        H.assert_parity(det['pre'][i - 1], g['prob_pre'][i - 1], 'pre-threshold day %d' % i)
        assert sparse.isspmatrix_coo(sol[i])
        H.assert_thresholded_parity(sol[i].toarray(), H.coo(g, 'prob%d' % i).toarray(), what='prob day %d' % i)
        assert abs(sol[i].sum() - 1) < H.MASS


def test_get_populations_small(pkb):
    z, days, pmfs = _small_pmfs()
    g = H.load('chain_small')
    n = 10
    rr, D, ms = int(z['rad_res']), int(g['dom_len']), g['max_shape']
    pop = pkb.CS.get_populations([H.recentre(pmfs[0], rr).tocsr()], pmfs[:n], days[:n], n, D, ms, 1, 130000,
                                 lambda d: 1.0)
    assert len(pop) == n
    for i in range(n):
        assert sparse.isspmatrix_csr(pop[i])
        H.assert_thresholded_parity(pop[i].toarray(), H.coo(g, 'pop1_%d' % i).toarray(), what='pop1 day %d' % i)
    r_dur = 3
    r_spread = [H.recentre(pmfs[i], rr).tocsr() for i in range(r_dur)]
    det = {}
    pop = pkb.CS.get_populations(r_spread, pmfs[:n], days[:n], n, D, ms, r_dur, 40000, lambda d: 1. / r_dur, details=det)
    for i in range(n):
        H.assert_thresholded_parity(pop[i].toarray(), H.coo(g, 'pop3_%d' % i).toarray(), what='pop3 day %d' % i)


# ---- fused solve (Run.main's hot path) ---------------------------------------
def _fused_small(pkb, tmp_path, **kw):
    z = H.load('pm_small')
    wind, days = pkb.PM.get_wind_data(H.write_wind_file(tmp_path, 'kalbar'), int(z['interp']), '00:00')
    n = 10
    w = pkb.Run.stack_wind(wind, days)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        res = pkb.Run.solve(w, n, H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, int(z['n_periods']), float(z['rad_dist']),
                            int(z['rad_res']), want_coo=True, want_dense=True, **kw)
    return res


def test_fused_solve_prob_small(pkb, tmp_path):
    g = H.load('chain_small')
    res = _fused_small(pkb, tmp_path, prob_model=True)
    assert res.ndays == 10 and res.dom_len == int(g['dom_len']) and res.max_shape == int(g['max_shape'][0])
    assert res.flags()[1:] == [bool(f) for f in g['prob_flags']]
    sols = res.coo_list()
    for i in range(10):
        ref = H.coo(g, 'prob%d' % i).toarray()
        H.assert_thresholded_parity(res.dense(i), ref, what='fused prob day %d' % i)
        H.assert_thresholded_parity(sols[i].toarray(), ref, what='fused prob COO day %d' % i)
        assert np.array_equal(sols[i].toarray(), res.dense(i))
        assert abs(sols[i].sum() - 1) < H.MASS
        # scipy.sparse.coo_matrix(dense) ordering: row-major
        order = np.lexsort((sols[i].col, sols[i].row))
        assert np.array_equal(order, np.arange(sols[i].nnz))
    cells = np.array([[40, 40], [0, 0], [41, 39], [80, 80]])
    smp = res.sample(cells)
    for i in range(10):
        assert np.array_equal(smp[i], res.dense(i)[cells[:, 0], cells[:, 1]])
    res.close()


def test_fused_solve_pop_small(pkb, tmp_path):
    g = H.load('chain_small')
    res = _fused_small(pkb, tmp_path, prob_model=False, r_dur=1, r_number=130000, r_dist=[1.0])
    for i in range(10):
        H.assert_thresholded_parity(res.dense(i), H.coo(g, 'pop1_%d' % i).toarray(), what='fused pop1 day %d' % i)
    res.close()
    res = _fused_small(pkb, tmp_path, prob_model=False, r_dur=3, r_number=40000, r_dist=[1. / 3] * 3)
    for i in range(10):
        H.assert_thresholded_parity(res.dense(i), H.coo(g, 'pop3_%d' % i).toarray(), what='fused pop3 day %d' % i)
    res.close()


def test_run_main_small(pkb, tmp_path, monkeypatch):
    """Params / main drop-in: output file format of Run.py:490-516."""
    monkeypatch.chdir(tmp_path)
    prefix = H.write_wind_file(tmp_path, 'kalbar')
    params = pkb.Run.Params()
    params.cmd_line_chg(['--kalbar', 'site_name=' + prefix, 'interp_num=2', 'n_periods=2', 'domain_info=(2000.0,40)',
                         'ndays=4', 'outfile=' + str(tmp_path / 'out' / 'run')])
    assert params.get_model_params() == (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, 2000.0, 40)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sol = pkb.Run.main(params)
    g = H.load('chain_small')
    assert len(sol) == 4
    for i in range(4):
        H.assert_thresholded_parity(sol[i].toarray(), H.coo(g, 'prob%d' % i).toarray(), what='main day %d' % i)
    saved = np.load(str(tmp_path / 'out' / 'run.npz'))
    assert list(saved['days']) == [13, 14, 15, 16]
    csr = sparse.csr_matrix((saved['14_data'], saved['14_ind'], saved['14_indptr']), shape=(81, 81))
    assert np.array_equal(csr.toarray(), sol[1].toarray())
    params.cmd_line_chg(['--pop'])
    assert params.PROB_MODEL is False and '_pop' in params.outfile


# ---- full-size configurations (GPU only) --------------------------------------
def _check_stats(g, prefix, sols, tol_nnz=3):
    for i, s in enumerate(sols):
        s = sparse.coo_matrix(s)
        assert abs(s.nnz - g[prefix + '_nnz'][i]) <= tol_nnz, (prefix, i, s.nnz, g[prefix + '_nnz'][i])
        scale = max(1.0, abs(g[prefix + '_sum'][i]))
        assert abs(s.data.sum() - g[prefix + '_sum'][i]) < 1e-10 * scale
        assert abs((s.data ** 2).sum() - g[prefix + '_sumsq'][i]) < 1e-10 * scale ** 2
        assert abs(s.data.max() - g[prefix + '_max'][i]) < 1e-10 * scale
        a = int(np.argmax(s.data))
        assert [s.row[a], s.col[a]] == list(g[prefix + '_argmax'][i])


@pytest.mark.gpu
@pytest.mark.parametrize('site', ['kalbar', 'carnarvon'])
def test_full_configs(gpu, tmp_path, site):
    """Configs 1-3: Kalbar / Carnarvon at default resolution, probability and
    population model, against statistics and full rows of the reference output."""
    g = H.load(site + '_full')
    wind, days = gpu.PM.get_wind_data(H.write_wind_file(tmp_path, site), 30, H.SITES[site])
    w = gpu.Run.stack_wind(wind, days)
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 30, 10000.0, 400)
    res = gpu.Run.solve(w, len(days), *args, prob_model=True, want_coo=True, want_dense=True)
    assert res.max_shape == int(g['max_shape'][0])
    assert res.flags()[1:] == [bool(f) for f in g['prob_flags']]
    sols = res.coo_list()
    _check_stats(g, 'prob', sols)
    for s in sols:
        assert abs(s.sum() - 1) < H.MASS
    last = res.dense(len(days) - 1)
    H.assert_thresholded_parity(last[400, :], g['prob_last_row400'], what='last day row 400')
    H.assert_thresholded_parity(last[:, 380], g['prob_last_col380'], what='last day col 380')
    res.close()
    r_dur, r_number, r_start = (1, 130000, None) if site == 'kalbar' else (5, 40000, 0.354)
    res = gpu.Run.solve(w, len(days), *args, prob_model=False, r_dur=r_dur, r_number=r_number,
                        r_dist=[1. / r_dur] * r_dur, r_start=r_start, want_coo=True, want_dense=True)
    assert res.max_shape == int(g['max_shape_pop'][0])
    _check_stats(g, 'pop', res.coo_list(), tol_nnz=6)
    last = res.dense(len(days) - 1)
    H.assert_thresholded_parity(last[400, :], g['pop_last_row400'], what='pop last day row 400', max_abs=1e-10)
    H.assert_thresholded_parity(last[:, 380], g['pop_last_col380'], what='pop last day col 380', max_abs=1e-10)
    res.close()


@pytest.mark.parametrize('n', [2, 3, 5, 7, 8, 12, 14, 30, 49, 64, 210, 360, 1000, 1764, 4704])
def test_debug_fft_matches_numpy(pkb, n):
    """The shared-memory FFT behind every chain kernel against numpy's pocketfft
    (the transform CalcSol.py:24,35 reaches through scipy.fftpack)."""
    import ctypes as C
    rng = np.random.default_rng(n)
    z = rng.normal(size=n) + 1j * rng.normal(size=n)
    zin = np.ascontiguousarray(z.view(np.float64))
    lib, ctx = pkb._lib.lib(), pkb._lib.ctx()
    for inverse, ref in ((0, np.fft.fft(z)), (1, np.fft.ifft(z) * n)):
        out = np.empty(2 * n)
        pkb._lib.check(lib.pkb_debug_fft(ctx.h, n, pkb._lib.dptr(zin), pkb._lib.dptr(out), inverse))
        got = out.view(np.complex128)
        assert np.abs(got - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()) * np.log2(n + 1)


def test_window_steps_match_whole_torus_steps(pkb):
    """While the state's support window fits inside the domain the fused solve runs
    each step on a torus sized for the window (ChainDims::win); the results must
    equal the whole-torus steps to rounding, for both models."""
    rng = np.random.default_rng(1)
    nd, periods, rad_res, rad_dist = 7, 96, 60, 3000.0
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        x = np.cumsum(rng.normal(0, 0.05, nd * periods)).reshape(nd, periods)
        w[:, :, c] = 0.3 * np.sin(np.linspace(0, 6, nd * periods)).reshape(nd, periods) + x * 0.2
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, rad_dist, rad_res)
    ctx = pkb._lib.ctx()
    for kw in (dict(prob_model=True), dict(prob_model=False, r_dur=1, r_number=1000.0)):
        out = {}
        for windows in (1, 0):
            ctx.set_option('windows', windows)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                res = pkb.Run.solve(w, nd, *args, want_coo=False, want_dense=True, **kw)
            out[windows] = ([res.dense(d) for d in range(nd)], res.flags(), res.N, res.window_steps())
            res.close()
        ctx.set_option('windows', 1)
        assert out[1][1] == out[0][1]
        assert out[1][3] > 0 and out[0][3] == 0, 'no step ran in window mode'
        for d in range(nd):
            a, b = out[1][0][d], out[0][0][d]
            assert ((a != 0) != (b != 0)).sum() == 0
            # population grids are probabilities scaled by r_number
            assert np.abs(a - b).max() <= 1e-15 * kw.get('r_number', 1.0)


def _oracle_solve(w, nd, args, rad_res):
    wind_data = {d: w[d] for d in range(w.shape[0])}     # the whole series: late take-offs of a day drift into the next
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pmfs = [PO.prob_mass(d, wind_data, *args) for d in range(nd)]
    D = 2 * rad_res + 1
    sol = [H.recentre(pmfs[0], rad_res)]
    if nd > 1:
        CO.get_solutions(sol, pmfs, list(range(nd)), nd, D, [max(p.shape[0] for p in pmfs)] * 2)
    return sol


def test_fused_solve_edge_cases(pkb):
    """Degenerate inputs of the fused solve against the oracle: a single day (no chain step at
    all), dead calm (every kernel is the local-diffusion blob only), and a gale that carries
    most take-offs out of the domain (loss bookkeeping + renormalisation, ParasitoidModel.py:546-558)."""
    nd, periods, rad_res, rad_dist = 3, 48, 30, 1500.0
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, rad_dist, rad_res)
    calm = np.zeros((nd, periods, 3))
    gale = np.zeros((nd, periods, 3))
    gale[:, :, 0] = 2.5
    gale[:, :, 1] = -1.0
    gale[:, :, 2] = np.hypot(gale[:, :, 0], gale[:, :, 1])
    breeze = np.zeros((nd, periods, 3))
    breeze[:, :, 0] = 0.3 * np.sin(np.linspace(0, 9, nd * periods)).reshape(nd, periods)
    breeze[:, :, 2] = np.abs(breeze[:, :, 0])
    for name, w, ndays in (('one day', breeze, 1), ('calm', calm, nd), ('gale', gale, nd)):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            res = pkb.Run.solve(w, ndays, *args, want_coo=True, want_dense=True)
        ref = _oracle_solve(w, ndays, args, rad_res)
        sols = res.coo_list()
        assert len(sols) == ndays
        for d in range(ndays):
            H.assert_thresholded_parity(res.dense(d), ref[d].toarray(), what='%s day %d' % (name, d))
            assert abs(sols[d].sum() - 1) < H.MASS
        res.close()
    with pytest.raises(pkb._lib.PkbError):
        pkb.Run.solve(breeze, nd + 1, *args)            # more days than wind
    with pytest.raises(pkb._lib.PkbError):
        pkb.Run.solve(breeze, nd, *args, prob_model=False, r_dur=nd + 1)


def test_fused_row_passes_match_separate_passes(pkb):
    """Option fuse_rows: the inverse row pass of a whole-torus step also runs the forward row
    transform of the next step on the row pair it still holds (k_rows_inv -> Yt), speculatively --
    a state that turns out flagged is transformed again from its truncated form.  Same arithmetic
    on the same numbers, so the solutions must agree to the run-to-run noise of kernel construction
    (its fp64 atomics commute only to an ulp) with and without it, flagged (steady breeze towards the
    edge) or not."""
    rng = np.random.default_rng(5)
    nd, periods, rad_res, rad_dist = 7, 96, 60, 3000.0
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        x = np.cumsum(rng.normal(0, 0.05, nd * periods)).reshape(nd, periods)
        w[:, :, c] = 0.3 * np.sin(np.linspace(0, 6, nd * periods)).reshape(nd, periods) + x * 0.2
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    drift = w.copy()
    drift[:, :, 0] = np.abs(drift[:, :, 0]) + 0.4       # mass reaches the boundary: flags trip
    drift[:, :, 2] = np.hypot(drift[:, :, 0], drift[:, :, 1])
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, rad_dist, rad_res)
    ctx = pkb._lib.ctx()
    ctx.set_option('windows', 0)                        # whole-torus steps from the first day on
    flagged = 0
    try:
        for wind in (w, drift):
            for kw in (dict(prob_model=True), dict(prob_model=False, r_dur=1, r_number=1000.0)):
                out = {}
                for fuse in (1, 0):
                    ctx.set_option('fuse_rows', fuse)
                    with warnings.catch_warnings():
                        warnings.simplefilter('ignore')
                        res = pkb.Run.solve(wind, nd, *args, want_coo=False, want_dense=True, **kw)
                    out[fuse] = ([res.dense(d) for d in range(nd)], res.flags())
                    res.close()
                assert out[1][1] == out[0][1]
                flagged += sum(1 for f in out[1][1] if f)
                for d in range(nd):
                    a, b = out[1][0][d], out[0][0][d]
                    assert ((a != 0) != (b != 0)).sum() == 0
                    assert np.abs(a - b).max() <= 1e-15 * kw.get('r_number', 1.0), 'day %d' % d
    finally:
        ctx.set_option('fuse_rows', 1)
        ctx.set_option('windows', 1)
    assert flagged > 0, 'no flagged step in the drift case'


def test_step_torus_matches_chain_torus(pkb):
    """Option step_torus: a whole-torus step runs on the smallest 7-smooth torus >= P + 2m of THAT day's
    kernel instead of the chain's (sized for the largest kernel); the fold mod P makes both the same
    circular convolution (CalcSol.py:66), so the solutions agree to rounding -- both models, flagged
    days included, against each other and (probability model) against the oracle."""
    rng = np.random.default_rng(9)
    nd, periods, rad_res, rad_dist = 7, 96, 60, 3000.0
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        x = np.cumsum(rng.normal(0, 0.05, nd * periods)).reshape(nd, periods)
        w[:, :, c] = 0.3 * np.sin(np.linspace(0, 6, nd * periods)).reshape(nd, periods) + x * 0.2
    w[:, :, 1] *= np.linspace(0.2, 2.0, nd)[:, None]          # kernel radii differ from day to day
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, rad_dist, rad_res)
    ctx = pkb._lib.ctx()
    ctx.set_option('windows', 0)
    ctx.set_option('spectral', 0)                       # (spectral-resident steps keep to the chain's torus)
    try:
        for kw in (dict(prob_model=True), dict(prob_model=False, r_dur=2, r_number=1000.0)):
            out = {}
            for st in (1, 0):
                ctx.set_option('step_torus', st)
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    res = pkb.Run.solve(w, nd, *args, want_coo=False, want_dense=True, **kw)
                out[st] = ([res.dense(d) for d in range(nd)], res.flags(), res.radii(), res.N, res.P)
                res.close()
            radii, N, P = out[1][2], out[1][3], out[1][4]
            lib = pkb._lib.lib()
            assert any(lib.pkb_smooth_len(P + 2 * m) < N for m in radii[1:]), 'no day would run on a smaller torus'
            assert out[1][1] == out[0][1]
            scale = kw.get('r_number', 1.0)
            for d in range(nd):
                H.assert_thresholded_parity(out[1][0][d] / scale, out[0][0][d] / scale, what='day %d' % d, max_abs=1e-15)
            if kw['prob_model']:
                ref = _oracle_solve(w, nd, args, rad_res)
                for d in range(nd):
                    H.assert_thresholded_parity(out[1][0][d], ref[d].toarray(), what='oracle day %d' % d)
    finally:
        ctx.set_option('step_torus', 1)
        ctx.set_option('windows', 1)
        ctx.set_option('spectral', 1)


def test_csr_output_equals_coo_output(pkb):
    """want_coo='csr': the same row-major triplets with per-row offsets instead of a row index per non-zero (12 instead of
    16 bytes each across PCIe; the layout Run.main saves, Run.py:490-510)."""
    rng = np.random.default_rng(2)
    nd, periods, rad_res, rad_dist = 5, 48, 30, 1500.0
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        w[:, :, c] = 0.25 * np.sin(np.linspace(0, 5 + c, nd * periods)).reshape(nd, periods) + rng.normal(0, 0.05, (nd, periods))
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, rad_dist, rad_res)
    for kw in (dict(prob_model=True), dict(prob_model=False, r_dur=2, r_number=1000.0)):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            a = pkb.Run.solve(w, nd, *args, want_coo=True, **kw)
            b = pkb.Run.solve(w, nd, *args, want_coo='csr', **kw)
        coo, csr = a.coo_list(), b.csr_list()
        for d in range(nd):
            assert sparse.isspmatrix_csr(csr[d]) and csr[d].nnz == coo[d].nnz
            assert np.array_equal(csr[d].toarray(), coo[d].toarray())
            assert csr[d].has_sorted_indices
        with pytest.raises(pkb._lib.PkbError):
            b.coo_arrays()
        a.close(); b.close()


def test_sparse_outputs_equal_dense_output(pkb):
    """COO / CSR output against the dense days of a separate solve, bit for bit.  With sparse output only, the emission
    writes and the compaction reads just the rows and columns a day's step computed (support windows, k_emit_dense
    sparse_only), and on the GPU a helper host thread drives the per-day compaction and copies (option coo_thread);
    neither may change a value or an index."""
    rng = np.random.default_rng(5)
    nd, periods, rad_res, rad_dist = 7, 48, 60, 3000.0
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        w[:, :, c] = 0.25 * np.sin(np.linspace(0, 5 + c, nd * periods)).reshape(nd, periods) + rng.normal(0, 0.05, (nd, periods))
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, 0.3 * H.MU_R, 2, rad_dist, rad_res)
    ctx = pkb._lib.ctx()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            dense = pkb.Run.solve(w, nd, *args, prob_model=True, want_coo=False, want_dense=True)
            assert dense.window_steps() >= 2          # the case the sparse-only emission is about
            ref = [dense.dense(d).copy() for d in range(nd)]
            dense.close()
            for thread in (1, 0):
                ctx.set_option('coo_thread', thread)
                a = pkb.Run.solve(w, nd, *args, prob_model=True, want_coo=True)
                b = pkb.Run.solve(w, nd, *args, prob_model=True, want_coo='csr')
                coo, csr = a.coo_list(), b.csr_list()
                for d in range(nd):
                    assert np.array_equal(coo[d].toarray(), ref[d]), (thread, d)
                    assert np.array_equal(csr[d].toarray(), ref[d]), (thread, d)
                    assert csr[d].nnz == np.count_nonzero(ref[d]) and csr[d].has_sorted_indices
                a.close(); b.close()
    finally:
        ctx.set_option('coo_thread', 1)


def test_cohort_lanes_equal_sequential_back_solves(pkb):
    """Population model with a release of several days: after the release the cohort back-solves and the emission of day n
    run on child contexts beside the main chain's step n + 1 (option cohort_lanes; on the GPU).  Same kernels on the same
    inputs in another order of launches: identical days and cohort flags."""
    rng = np.random.default_rng(9)
    nd, periods, rad_res, rad_dist = 8, 48, 40, 2000.0
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        w[:, :, c] = 0.3 * np.sin(np.linspace(0, 6 + c, nd * periods)).reshape(nd, periods) + rng.normal(0, 0.05, (nd, periods))
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, rad_dist, rad_res)
    kw = dict(prob_model=False, r_dur=3, r_number=5000.0, r_start=0.3)
    ctx = pkb._lib.ctx()
    out = {}
    try:
        for on in (1, 0):
            ctx.set_option('cohort_lanes', on)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                res = pkb.Run.solve(w, nd, *args, want_coo='csr', want_dense=True, **kw)
            out[on] = ([res.dense(d).copy() for d in range(nd)], [m.toarray() for m in res.csr_list()], list(res.flags()),
                       [list(res.cohort_flags(d, 2)) for d in range(3, nd)])
            res.close()
    finally:
        ctx.set_option('cohort_lanes', 1)
    for d in range(nd):
        assert np.array_equal(out[1][0][d], out[0][0][d]), d
        assert np.array_equal(out[1][1][d], out[1][0][d]), d          # the CSR output is the dense day
    assert out[1][2] == out[0][2] and out[1][3] == out[0][3]


def test_spectral_steps_match_exact_steps(pkb):
    """Option spectral: while the content outside the domain is below 1e-13 the chain keeps the product spectrum
    k_cols forms anyway and starts the next step from it (the reference's own chain state is spectral,
    CalcSol.py:66,189-201) instead of re-transforming the folded real state.  The device decides per step; the
    bound on the difference is 1e-12 (chain.cuh).  Calm wind: the mass stays inside, spectral steps must happen
    and agree with the exact steps and the oracle.  Breeze towards the edge: the flags trip, the chain must fall
    back to exact steps and give the same flags and solutions."""
    rng = np.random.default_rng(5)
    nd, periods, rad_res, rad_dist = 10, 96, 140, 7000.0
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        x = np.cumsum(rng.normal(0, 0.05, nd * periods)).reshape(nd, periods)
        w[:, :, c] = 0.15 * np.sin(np.linspace(0, 6, nd * periods)).reshape(nd, periods) + x * 0.1
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    drift = w.copy()
    drift[:, :, 0] = np.abs(drift[:, :, 0]) * 2 + 0.8
    drift[:, :, 2] = np.hypot(drift[:, :, 0], drift[:, :, 1])
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, rad_dist, rad_res)
    ctx = pkb._lib.ctx()
    seen_spec = seen_flag = seen_rows = 0
    try:
        for windows in (0, 1):
            ctx.set_option('windows', windows)
            for wind in (w, drift):
                for kw in (dict(prob_model=True), dict(prob_model=False, r_dur=1, r_number=1000.0)):
                    out = {}
                    for sp in (1, 0):
                        ctx.set_option('spectral', sp)
                        with warnings.catch_warnings():
                            warnings.simplefilter('ignore')
                            res = pkb.Run.solve(wind, nd, *args, want_coo=False, want_dense=True, keep_pre=True, **kw)
                        out[sp] = ([res.dense(d) for d in range(nd)], res.flags(), res.spectral_steps(), [res.pre(d) for d in range(nd)],
                                   res.row_windows())
                        res.close()
                    assert out[0][2] == []
                    assert out[1][1] == out[0][1]
                    seen_rows += sum(1 for w_ in out[1][4] if w_ is not None)
                    # a spectral step never follows a flagged state
                    assert not any(out[1][1][d - 1] for d in out[1][2])
                    seen_spec += len(out[1][2])
                    seen_flag += sum(1 for f in out[1][1] if f)
                    scale = kw.get('r_number', 1.0)
                    for d in range(nd):
                        assert np.abs(out[1][3][d] - out[0][3][d]).max() <= 1e-13 * scale, 'pre-threshold day %d' % d
                        H.assert_thresholded_parity(out[1][0][d] / scale, out[0][0][d] / scale, what='day %d' % d, max_abs=1e-13)
                    if kw['prob_model']:
                        ref = _oracle_solve(wind, nd, args, rad_res)
                        for d in range(nd):
                            H.assert_thresholded_parity(out[1][0][d], ref[d].toarray(), what='oracle day %d' % d)
    finally:
        ctx.set_option('spectral', 1)
        ctx.set_option('windows', 1)
    assert seen_spec >= 8, 'no spectral-resident steps in the calm case'
    assert seen_rows >= 4, 'no row-windowed spectral step (option spectral_rows) in the calm case'
    assert seen_flag > 0, 'no flagged step in the drift case'


def test_trunc_torus_matches_full_torus(pkb):
    """Option trunc_torus: a step whose source state was truncated by the boundary flag (CalcSol.py:200-201)
    spans only D + 2m cells, so it runs on a torus >= D + 2m instead of >= P + 2m; the kernels pick the
    geometry on the device from the flag.  Same circular convolution mod P, so the solutions agree to
    rounding with the option off and with the oracle -- both models, with flags tripping."""
    rng = np.random.default_rng(5)
    nd, periods, rad_res, rad_dist = 7, 96, 60, 3000.0
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        x = np.cumsum(rng.normal(0, 0.05, nd * periods)).reshape(nd, periods)
        w[:, :, c] = 0.3 * np.sin(np.linspace(0, 6, nd * periods)).reshape(nd, periods) + x * 0.2
    w[:, :, 0] = np.abs(w[:, :, 0]) + 0.4               # mass reaches the boundary: flags trip
    w[:, :, 1] *= np.linspace(0.2, 2.0, nd)[:, None]    # kernel radii differ from day to day
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    args = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, rad_dist, rad_res)
    ctx = pkb._lib.ctx()
    lib = pkb._lib.lib()
    try:
        for windows in (0, 1):
            ctx.set_option('windows', windows)
            for kw in (dict(prob_model=True), dict(prob_model=False, r_dur=1, r_number=1000.0), dict(prob_model=False, r_dur=2, r_number=1000.0)):
                out = {}
                for tt in (1, 0):
                    ctx.set_option('trunc_torus', tt)
                    with warnings.catch_warnings():
                        warnings.simplefilter('ignore')
                        res = pkb.Run.solve(w, nd, *args, want_coo=False, want_dense=True, **kw)
                    out[tt] = ([res.dense(d) for d in range(nd)], res.flags(), res.radii(), res.dom_len, res.P)
                    res.close()
                flags, radii, D, P = out[1][1:]
                assert out[0][1] == flags
                # a flagged state is followed by a step whose truncated-source torus is smaller than its own
                assert any(flags[n - 1] and lib.pkb_smooth_len(D + 2 * radii[n]) < lib.pkb_smooth_len(P + 2 * radii[n])
                           for n in range(2, nd)), 'no step could have used the smaller torus'
                scale = kw.get('r_number', 1.0)
                for d in range(nd):
                    H.assert_thresholded_parity(out[1][0][d] / scale, out[0][0][d] / scale, what='day %d' % d, max_abs=1e-15)
                if kw['prob_model']:
                    ref = _oracle_solve(w, nd, args, rad_res)
                    for d in range(nd):
                        H.assert_thresholded_parity(out[1][0][d], ref[d].toarray(), what='oracle day %d' % d)
    finally:
        ctx.set_option('trunc_torus', 1)
        ctx.set_option('windows', 1)
