"""Batched-likelihood path (parasitoids_b200/batch.py): sharding of proposals
over ranks and the one all_gather of the sampled cells.  The world_size-2 test
runs on CPU with the gloo backend; each rank drives the emulated build of the
kernels (tests/emul, test infrastructure) so that the whole host path --
partition, per-proposal fused solve, gather, re-ordering -- is exercised."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DEFAULT = np.array([1.263, 3.913, 7.302, 2.614, 23.999, 2.350, 171.82, 144.58, 0.253, 7.096, 7.260, 0.0, 1.0, 2, 1.179])


def _wind(nd=4, periods=48, seed=3):
    rng = np.random.default_rng(seed)
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        w[:, :, c] = 0.25 * np.sin(np.linspace(0, 5 + c, nd * periods)).reshape(nd, periods) + rng.normal(0, 0.05, (nd, periods))
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    return w


def _proposals(B, seed=11):
    rng = np.random.default_rng(seed)
    P = np.tile(DEFAULT, (B, 1))
    P[:, 6] *= 1 + 0.2 * rng.uniform(-1, 1, B)      # sig_x
    P[:, 7] *= 1 + 0.2 * rng.uniform(-1, 1, B)      # sig_y
    P[:, 8] = rng.uniform(-0.4, 0.4, B)             # corr
    P[:, 13] = rng.integers(1, 4, B)                # n_periods
    P[:, 14] *= 1 + 0.3 * rng.uniform(-1, 1, B)     # mu_r
    return P


CELLS = np.array([[20, 20], [18, 23], [25, 14], [0, 0], [40, 40], [20, 30]], dtype=np.int32)
SOLVE_KW = dict(ndays=4, rad_dist=1000.0, rad_res=20, prob_model=False, r_dur=2, r_number=1000.0, r_start=0.3)


def test_shard_partitions_every_item_once():
    from parasitoids_b200 import batch
    for n in (0, 1, 5, 8, 512, 513):
        for world in (1, 2, 3, 8):
            parts = [batch.shard(n, world, r) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        batch.shard(4, 2, 2)
    # cost-aware dealing: still a partition into equal counts, and the shares cost about the same
    rng = np.random.default_rng(0)
    for n in (5, 64, 512):
        cost = rng.gamma(2.0, 1.0, n)
        for world in (2, 3, 8):
            parts = [batch.shard(n, world, r, cost) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            if n == 512:
                tot = [cost[p].sum() for p in parts]
                assert max(tot) - min(tot) < 0.02 * np.mean(tot)


def test_unpack_proposal_order():
    from parasitoids_b200 import batch
    hp, dp, dl, mu_r, n_periods = batch.unpack_proposal(DEFAULT)
    # Run.py:374-379: hparams = (lam, *g_params, *f_params)
    assert hp == (1.0, 1.263, 3.913, 7.302, 2.614, 23.999, 2.350)
    assert dp == (171.82, 144.58, 0.253) and dl == (7.096, 7.260, 0.0)
    assert mu_r == 1.179 and n_periods == 2
    with pytest.raises(ValueError):
        batch.unpack_proposal(DEFAULT[:-1])


@pytest.mark.parametrize('prob_model,lanes,group', [(False, 2, 32), (False, 1, 2), (True, 3, 2)])
def test_solve_batch_matches_single_solves(pkb, prob_model, lanes, group):
    """pkb_solve_batch (kernel construction batched over proposals, chains enqueued round-robin on
    `batch_lanes` child contexts, sample-cell emission only) against one Run.solve per proposal
    sampled at the same cells of its dense output, for the population and the probability model;
    group = 2 splits the five proposals into three pipelined kernel-construction groups."""
    import warnings
    from parasitoids_b200 import batch
    w, props = _wind(), _proposals(5)
    kw = dict(SOLVE_KW)
    if prob_model:
        kw.update(prob_model=True, r_dur=1, r_number=1.0, r_start=None)
    ctx = pkb._lib.ctx()
    ctx.set_option('batch_lanes', lanes)
    ctx.set_option('batch_group', group)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            l0 = ctx.launch_count()
            got = batch.solve_batch(w, props, CELLS, **kw)
            assert ctx.launch_count() - l0 > 5 * kw['ndays']      # the lanes' launches are folded into the parent's count
            for b in range(5):
                hp, dp, dl, mu_r, n_periods = batch.unpack_proposal(props[b])
                res = pkb.Run.solve(w, kw['ndays'], hp, dp, dl, mu_r, n_periods, kw['rad_dist'], kw['rad_res'],
                                    prob_model=prob_model, r_dur=kw['r_dur'], r_number=kw['r_number'], r_start=kw['r_start'],
                                    want_coo=False, keep_device=True)
                ref = res.sample(CELLS)
                res.close()
                assert ((got[b] != 0) != (ref != 0)).sum() == 0
                assert np.allclose(got[b], ref, rtol=1e-12, atol=1e-15)
    finally:
        ctx.set_option('batch_lanes', 4)
        ctx.set_option('batch_group', 32)


@pytest.mark.parametrize('prob_model,group,sprd,big', [(False, 32, None, False), (False, 2, None, True), (True, 32, None, True),
                                                       (False, 32, 0.6, False)])
def test_batched_chain_matches_per_proposal_chain(pkb, prob_model, group, sprd, big):
    """Batched chain kernels (csrc/bchain.cuh: step n of every proposal of a group in one launch per pass) against
    (i) the per-proposal chains of the same library call (option batch_chain = 0) and (ii) one Run.solve per proposal,
    for the population model with a one-day release (the Kalbar setting of Bayes_Run.py), the probability model and
    the leading spread day.  The launch count shows which path ran.  Small domain: every step flagged (truncated-source
    torus) and one proposal with stencil-sized kernels (FFT steps on the batched path); big domain:
    support-window steps, un-flagged and flagged whole-torus steps side by side in one launch."""
    import warnings
    from parasitoids_b200 import batch
    w, props = _wind(nd=8 if big else 6), _proposals(5)
    props[1, 14] *= 2.5                     # one plume that reaches the boundary early
    props[1, 13] = 3
    kw = dict(SOLVE_KW, ndays=6, r_dur=1, r_start=None)
    if big:
        props[:, 14] *= 0.25
        kw.update(ndays=8, rad_dist=3000.0, rad_res=60)
    if prob_model:
        kw.update(prob_model=True, r_number=1.0)
    if sprd is not None:
        kw.update(ndays=5, sprd_factor=sprd)
    ctx = pkb._lib.ctx()
    ctx.set_option('batch_group', group)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            l0 = ctx.launch_count()
            got = batch.solve_batch(w, props, CELLS, **kw)
            n_batched = ctx.launch_count() - l0
            ctx.set_option('batch_chain', 0)
            l0 = ctx.launch_count()
            ref = batch.solve_batch(w, props, CELLS, **kw)
            n_single = ctx.launch_count() - l0
            assert n_batched < n_single, (n_batched, n_single)
            assert ((got != 0) != (ref != 0)).sum() == 0
            assert np.allclose(got, ref, rtol=1e-12, atol=1e-15)
            if not prob_model:
                # same steps, same jobs, same summation orders: the population path (no spectral-resident steps, no tau windows) is
                # bit-identical -- except for a proposal with stencil-sized kernels, which the batched path runs through the FFT
                assert np.array_equal(np.delete(got, 1, axis=0), np.delete(ref, 1, axis=0))
            if sprd is None:
                for b in range(5):
                    hp, dp, dl, mu_r, n_periods = batch.unpack_proposal(props[b])
                    res = pkb.Run.solve(w, kw['ndays'], hp, dp, dl, mu_r, n_periods, kw['rad_dist'], kw['rad_res'],
                                        prob_model=prob_model, r_dur=kw['r_dur'], r_number=kw['r_number'], r_start=kw['r_start'],
                                        want_coo=False, keep_device=True)
                    one = res.sample(CELLS)
                    res.close()
                    assert ((got[b] != 0) != (one != 0)).sum() == 0
                    assert np.allclose(got[b], one, rtol=1e-12, atol=1e-15)
    finally:
        ctx.set_option('batch_chain', 1)
        ctx.set_option('batch_group', 32)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, outdir, B):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), PKB_EMUL_THREADS='2')
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import warnings
    import torch.distributed as dist
    import conftest
    conftest._activate('emul')
    from parasitoids_b200 import batch
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            out = batch.solve_batch(_wind(), _proposals(B), CELLS, **SOLVE_KW)
        np.save(os.path.join(outdir, 'rank%d.npy' % rank), out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('B', [3, 4])
def test_solve_batch_world2_gloo_matches_serial(pkb, tmp_path, B):
    if pkb.is_gpu:
        pytest.skip('host-side sharding logic: covered by the CPU run')
    import warnings
    import torch.multiprocessing as mp
    from parasitoids_b200 import batch
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        serial = batch.solve_batch(_wind(), _proposals(B), CELLS, **SOLVE_KW)
    assert serial.shape == (B, 4, len(CELLS))
    assert np.isfinite(serial).all() and serial[:, :, 0].min() > 0      # release cell is populated
    mp.spawn(_rank_main, args=(2, _free_port(), str(tmp_path), B), nprocs=2, join=True)
    r0 = np.load(tmp_path / 'rank0.npy')
    r1 = np.load(tmp_path / 'rank1.npy')
    assert np.array_equal(r0, r1), 'ranks disagree after the all_gather'
    # phase 1 accumulates with floating-point atomics, so two runs agree to rounding, not bitwise
    assert np.allclose(r0, serial, rtol=1e-12, atol=1e-15), 'sharded result differs from the serial loop'


@pytest.mark.gpu
def test_solve_batch_gpu_matches_oracle(gpu):
    """One GPU, three proposals: sampled cells against the oracle's population model."""
    import warnings
    from scipy import sparse
    from oracle import pm_oracle as PO, cs_oracle as CO
    from parasitoids_b200 import batch
    w = _wind()
    props = _proposals(3)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        got = batch.solve_batch(w, props, CELLS, **SOLVE_KW)
        nd, rr, rd = SOLVE_KW['ndays'], SOLVE_KW['rad_res'], SOLVE_KW['r_dur']
        D = 2 * rr + 1
        wind_data = {d: w[d] for d in range(nd)}
        for b in range(3):
            hp, dp, dl, mu_r, n_periods = batch.unpack_proposal(props[b])
            pmfs = [PO.prob_mass(d, wind_data, hp, dp, dl, mu_r, n_periods, SOLVE_KW['rad_dist'], rr,
                                 SOLVE_KW['r_start'] if d == 0 else None) for d in range(nd)]
            ms = [max(p.shape[0] for p in pmfs)] * 2
            r_spread = []
            for p in pmfs[:rd]:
                off = rr - p.shape[0] // 2
                r_spread.append(sparse.coo_matrix((p.data, (p.row + off, p.col + off)), shape=(D, D)).tocsr())
            pop = CO.get_populations(r_spread, pmfs, list(range(nd)), nd, D, ms, rd, SOLVE_KW['r_number'], lambda day: 1.0 / rd)
            ref = np.array([[pop[d][r, c] for r, c in CELLS] for d in range(nd)])
            both = (got[b] != 0) & (ref != 0)
            assert np.abs(np.where(both, got[b] - ref, 0)).max() <= 1e-10 * max(1.0, np.abs(ref).max())
            assert ((got[b] != 0) != (ref != 0)).sum() == 0
