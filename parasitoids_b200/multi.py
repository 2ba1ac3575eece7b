"""One forward solve spread over the GPUs of a box, one process per GPU.

The reference parallelises a single run only in phase 1 -- ``Pool.starmap(prob_mass)``
hands one day to each worker (Run.py:412-425) -- and runs the chain on one core.  Here
both phases shard:

phase 1   days round-robin over the ranks; every rank builds its kernels on its GPU,
          the thresholded kernels (a few hundred cells a side) are all-gathered once.
phase 2   the chain state lives in Fourier space, as the reference's does
          (CalcSol.py:66,189-201), sharded by spectral column; a day is a column pass
          on the own columns, ONE all-to-all of the column-transformed product, and a
          row pass on the own rows (csrc/dist.cuh).  The per-day sums of the
          renormalisation travel in one 4-double all-gather.

That chain skips the fold mod P, which is only legitimate while nothing above 1e-13
lies outside the domain (the criterion of csrc/chain.cuh, here checked on the global
maxima of every day).  When it fails -- the population reaches the boundary -- the
solve is simply repeated on every rank with the exact single-GPU chain: a flagged
chain is sequential in the day index, it does not shard ("replicas only").

Collectives go through ``torch.distributed`` (NCCL over NVLink on GPUs; the CPU tests
use gloo with the emulated kernels).
"""
import ctypes as C

import numpy as np

from . import _abi, _lib, Run
from . import ParasitoidModel as PM


class MultiResult(object):
    """This rank's rows of every day's solution.

    rows      (ndays, rows_per_rank, dom_len) torch tensor on the solve's device; row r of the slab is domain row
              ``rank * rows_per_rank + r`` (rows beyond the domain are zero)
    sharded   False if the solve fell back to the replicated exact chain
    """

    def __init__(self, rows, rank, world, dom_len, ndays, meta, sharded, P, N):
        self.rows, self.rank, self.world, self.dom_len, self.ndays = rows, rank, world, dom_len, ndays
        self.meta, self.sharded, self.P, self.N = meta, sharded, P, N

    def gather(self, group=None):
        """Full (ndays, dom_len, dom_len) ndarray on every rank (tests and small runs: ndays * dom_len^2 doubles)."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return self.rows[:, :self.dom_len].cpu().numpy()
        out = torch.empty((self.world * self.rows.shape[0],) + tuple(self.rows.shape[1:]), dtype=self.rows.dtype, device=self.rows.device)
        dist.all_gather_into_tensor(out, self.rows.contiguous(), group=group)      # concatenated along dim 0
        out = out.view((self.world,) + tuple(self.rows.shape))
        full = out.permute(1, 0, 2, 3).reshape(self.ndays, -1, self.dom_len)[:, :self.dom_len]
        return full.cpu().numpy()


def _fallback(wind, ndays, model, rank, world, per, torch, dev):
    """Replicated exact solve; returns this rank's row slab in the same layout."""
    res = Run.solve(wind, ndays, *model, prob_model=True, want_coo=False, want_dense=True)
    try:
        D = res.dom_len
        rows = torch.zeros((ndays, per, D), dtype=torch.float64, device=dev)
        r0, r1 = min(D, rank * per), min(D, (rank + 1) * per)
        if r1 > r0:
            for d in range(ndays):
                rows[d, :r1 - r0] = torch.from_numpy(res.dense(d)[r0:r1]).to(dev)
        return rows, res.P, res.N
    finally:
        res.close()


def solve_single(wind, ndays, hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res, group=None, device=None,
                 allow_fallback=True):
    """Probability model of ONE run (CalcSol.get_solutions, Run.py:399-465) over all ranks of ``group``.

    wind: (nd_wind, periods, 3) ndarray, identical on every rank.  Returns a ``MultiResult``."""
    import torch
    import torch.distributed as dist
    model = (hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res)
    have = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if have else 1
    rank = dist.get_rank(group) if have else 0
    on_gpu = torch.cuda.is_available() and (not have or dist.get_backend(group) == 'nccl')
    dev = torch.device('cuda', device if device is not None else torch.cuda.current_device()) if on_gpu else torch.device('cpu')
    lib, ctx = _lib.lib(), _lib.ctx(device)
    D = 2 * int(rad_res) + 1
    wind = _lib.as_f64(wind)

    # ---- phase 1: my days ---------------------------------------------------------------------------
    mine = list(range(rank, ndays, world))
    nper = -(-ndays // world)
    rads = torch.zeros(nper, dtype=torch.int32)
    ks = None
    if mine:
        args = [PM._day_args(hparams, Dparams, Dlparams, mu_r, n_periods, rad_dist, rad_res, None, d, False) for d in mine]
        ks = PM.build_kernels(wind, args)
        for i, d in enumerate(mine):
            meta = ks.meta(i)
            PM._raise_for_status(meta, d, model)
            rads[i] = meta.rad
    rads = rads.to(dev)
    allr = torch.empty(world * nper, dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(allr, rads, group=group)
    else:
        allr.copy_(rads)
    allr_h = allr.cpu().numpy().reshape(world, nper)
    W = 2 * int(allr_h.max()) + 1
    local = torch.zeros((nper, W, W), dtype=torch.float64, device=dev)
    if on_gpu:
        _lib.check(lib.pkb_set_stream(ctx.h, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    try:
        for i in range(len(mine)):
            _lib.check(lib.pkb_kset_export_device(ks.h, i, C.c_void_p(local[i].data_ptr()), W))
        if ks is not None:
            ctx.sync()
            ks.close()
        allk = torch.empty((world * nper, W, W), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_gather_into_tensor(allk, local, group=group)
        else:
            allk.copy_(local)
        order = torch.tensor([(d % world) * nper + d // world for d in range(ndays)], device=dev)
        kernels = allk.index_select(0, order).contiguous()
        radii = np.ascontiguousarray([allr_h[d % world, d // world] for d in range(ndays)], dtype=np.int32)
        del allk, local
        kh = C.c_void_p()
        _lib.check(lib.pkb_kset_from_device(ctx.h, C.c_void_p(kernels.data_ptr()), ndays, W, _lib.iptr(radii), int(rad_res), C.byref(kh)))
        kset = PM.KernelSet(kh, ndays)
        # ---- phase 2 --------------------------------------------------------------------------------
        try:
            xe, per, P, N = C.c_longlong(), C.c_int(), C.c_int(), C.c_int()
            _lib.check(lib.pkb_dist_plan(ctx.h, D, int(radii.max()), world, C.byref(xe), C.byref(per), C.byref(P), C.byref(N)))
            send = torch.empty(2 * xe.value, dtype=torch.float64, device=dev)
            recv = torch.empty(2 * xe.value, dtype=torch.float64, device=dev)
            stats = torch.zeros(4, dtype=torch.float64, device=dev)
            allstats = torch.zeros(4 * world, dtype=torch.float64, device=dev)
            rows = torch.empty((ndays, per.value, D), dtype=torch.float64, device=dev)
            h = C.c_void_p()
            _lib.check(lib.pkb_dist_create(ctx.h, kset.h, ndays, rank, world, C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()),
                                           C.c_void_p(stats.data_ptr()), C.c_void_p(allstats.data_ptr()), C.c_void_p(rows.data_ptr()), C.byref(h)))
            try:
                for day in range(1, ndays):
                    _lib.check(lib.pkb_dist_step_cols(h, day))
                    if world > 1:
                        dist.all_to_all_single(recv, send, group=group)
                    else:
                        recv.copy_(send)
                    _lib.check(lib.pkb_dist_step_rows(h))
                    if world > 1:
                        dist.all_gather_into_tensor(allstats, stats, group=group)
                    else:
                        allstats.copy_(stats)
                    _lib.check(lib.pkb_dist_step_emit(h, day))
                meta = np.zeros((ndays, 4))
                ok = C.c_int()
                _lib.check(lib.pkb_dist_finish(h, _lib.dptr(meta), C.byref(ok)))
            finally:
                lib.pkb_dist_destroy(h)
        finally:
            kset.close()
    finally:
        if on_gpu:
            lib.pkb_set_stream(ctx.h, None)
    if ok.value:
        return MultiResult(rows, rank, world, D, ndays, meta, True, P.value, N.value)
    if not allow_fallback:
        raise RuntimeError('the population reaches the domain boundary (max outside the domain {:.3e}): this solve does not '
                           'shard over GPUs'.format(float(meta[:, 3].max())))
    rows, Pf, Nf = _fallback(wind, ndays, model, rank, world, per.value, torch, dev)
    return MultiResult(rows, rank, world, D, ndays, meta, False, Pf, Nf)
