"""One solve over several ranks (parasitoids_b200/multi.py, csrc/dist.cuh): days of phase 1 round-robin over the
ranks, the spectral-resident chain sharded by spectral column / row with one all-to-all per day.  The world_size-2
tests run on CPU with the gloo backend; each rank drives the emulated build of the kernels (tests/emul)."""
import os
import socket
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import helpers as H        # noqa: E402


def _wind(nd, calm=True, periods=96, seed=5):
    rng = np.random.default_rng(seed)
    w = np.zeros((nd, periods, 3))
    for c in range(2):
        x = np.cumsum(rng.normal(0, 0.05, nd * periods)).reshape(nd, periods)
        w[:, :, c] = 0.15 * np.sin(np.linspace(0, 6, nd * periods)).reshape(nd, periods) + x * 0.1
    if not calm:
        w[:, :, 0] = np.abs(w[:, :, 0]) * 2 + 0.8          # steady breeze: the population reaches the boundary
    w[:, :, 2] = np.hypot(w[:, :, 0], w[:, :, 1])
    return w


ND, RAD_RES, RAD_DIST = 6, 100, 5000.0
ARGS = (H.HPARAMS, H.DPARAMS, H.DLPARAMS, H.MU_R, 2, RAD_DIST, RAD_RES)


def _compare(full, ref, nd, tol=1e-15):
    for d in range(nd):
        a, b = full[d], ref[d]
        assert ((a != 0) != (b != 0)).sum() == 0, 'support of day %d' % d
        assert np.abs(a - b).max() <= tol, 'day %d' % d
        assert abs(a.sum() - 1) < H.MASS


def _reference(pkbRun, w):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        res = pkbRun.solve(w, ND, *ARGS, want_coo=False, want_dense=True)
    try:
        return [res.dense(d) for d in range(ND)], res.flags()
    finally:
        res.close()


def test_single_rank_matches_fused_solve(pkb):
    """world = 1: the slab-decomposed chain (no fold, spectral from day 1) against the fused exact solve."""
    from parasitoids_b200 import multi
    w = _wind(ND)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        r = multi.solve_single(w, ND, *ARGS, allow_fallback=False)
    assert r.sharded and float(r.meta[:, 3].max()) < 1e-13
    ref, flags = _reference(pkb.Run, w)
    assert not any(flags)
    _compare(r.gather(), ref, ND)


def test_boundary_case_falls_back(pkb):
    """When the population reaches the boundary the fold mod P matters: the sharded chain must notice (on the
    global per-day maxima) and the solve must come back from the exact chain."""
    from parasitoids_b200 import multi
    w = _wind(ND, calm=False)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        with pytest.raises(RuntimeError):
            multi.solve_single(w, ND, *ARGS, allow_fallback=False)
        r = multi.solve_single(w, ND, *ARGS)
    assert not r.sharded
    ref, flags = _reference(pkb.Run, w)
    assert any(flags)
    _compare(r.gather(), ref, ND, tol=0.0)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, outdir, calm):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), PKB_EMUL_THREADS='2')
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    import conftest
    conftest._activate('emul')
    from parasitoids_b200 import multi
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            r = multi.solve_single(_wind(ND, calm), ND, *ARGS)
            full = r.gather()
        np.save(os.path.join(outdir, 'rank%d.npy' % rank), full)
        np.save(os.path.join(outdir, 'sharded%d.npy' % rank), np.array([int(r.sharded), r.rows.shape[1]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,calm', [(2, True), (3, True), (2, False)])
def test_world_gloo_matches_single_gpu_solve(pkb, tmp_path, world, calm):
    if pkb.is_gpu:
        pytest.skip('host-side sharding logic: covered by the CPU run (the GPU run of this path is bench.py --single-solve)')
    import torch.multiprocessing as mp
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path), calm), nprocs=world, join=True)
    ref, flags = _reference(pkb.Run, _wind(ND, calm))
    outs = [np.load(tmp_path / ('rank%d.npy' % r)) for r in range(world)]
    for o in outs[1:]:
        assert np.array_equal(o, outs[0]), 'ranks disagree after the gather'
    sharded = [int(np.load(tmp_path / ('sharded%d.npy' % r))[0]) for r in range(world)]
    assert sharded == [1 if calm else 0] * world
    assert any(flags) == (not calm)
    _compare(outs[0], ref, ND, tol=1e-15 if calm else 0.0)
