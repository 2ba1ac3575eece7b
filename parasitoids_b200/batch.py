"""Batched likelihood path: many independent parameter proposals, each a
complete forward solve, sharded over the GPUs of one box.

The reference evaluates ONE proposal at a time: PyMC's AdaptiveMetropolis
writes the 15 block variables into ``Params`` and calls ``pop_model``
(Bayes_Run.py:186-196, 204-336), which runs ``prob_mass`` for every day and
then ``get_populations``; the likelihood only reads the model at the sample
cells (Bayes_funcs.py:58-74, 167-173).  A batch of B proposals is therefore B
independent units of work: they are partitioned over the ranks with no
data-path collective, every rank solves its share on its own GPU, and the
sampled-cell values -- the only thing the likelihood needs from every proposal
-- are exchanged with ONE all_gather (NCCL over NVLink when the process group
is NCCL; the CPU tests use gloo).
"""
import numpy as np

from . import Run

# order of the block-updated variables (Bayes_Run.py:186-187, 218-231)
PROPOSAL_FIELDS = ('g_aw', 'g_bw', 'f_a1', 'f_b1', 'f_a2', 'f_b2', 'sig_x', 'sig_y', 'corr',
                   'sig_x_l', 'sig_y_l', 'corr_l', 'lam', 'n_periods', 'mu_r')


def unpack_proposal(p):
    """15 block variables -> (hparams, Dparams, Dlparams, mu_r, n_periods) in the
    argument order of ``prob_mass`` (Run.py:374-379, Bayes_Run.py:218-231)."""
    p = np.asarray(p, dtype=float).ravel()
    if p.size != len(PROPOSAL_FIELDS):
        raise ValueError('a proposal has {} entries, got {}'.format(len(PROPOSAL_FIELDS), p.size))
    hparams = (p[12], p[0], p[1], p[2], p[3], p[4], p[5])
    return hparams, tuple(p[6:9]), tuple(p[9:12]), float(p[14]), int(round(p[13]))


def shard(n_items, world, rank, cost=None):
    """Indices of the items rank ``rank`` of ``world`` owns.  Without ``cost``: contiguous blocks, the first
    ``n_items % world`` ranks take one extra item.  With ``cost`` (one number per item): the items are dealt out in
    order of decreasing cost, back and forth over the ranks, so that every rank gets the same number of items (+-1) and
    about the same total cost -- the ranks of a likelihood batch are timed by the slowest one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError('bad rank {} of {}'.format(rank, world))
    n_items = int(n_items)
    if cost is not None:
        order = np.argsort(-np.asarray(cost, dtype=float), kind='stable')
        mine = []
        for pos, item in enumerate(order):
            rnd, k = divmod(pos, world)
            if (k if rnd % 2 == 0 else world - 1 - k) == rank:
                mine.append(int(item))
        return sorted(mine)
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def proposal_cost(proposals):
    """Relative cost of a proposal's solve, known before anything is computed: the kernel radii -- and with them the
    torus every chain step runs on -- grow with the flight advection scale mu_r * n_periods and the in-flow spread."""
    p = np.asarray(proposals, dtype=float).reshape(-1, len(PROPOSAL_FIELDS))
    return p[:, 14] * p[:, 13] + 0.02 * np.maximum(p[:, 6], p[:, 7])


_PINNED = {}
_COPY_OUT = False      # the returned array is a view of a cached pinned buffer: valid until the next solve_batch of that shape


def _pinned(shape):
    """Cached pinned host tensor for the gathered result (allocating pinned memory costs more than the copy)."""
    import torch
    if shape not in _PINNED:
        _PINNED.clear()
        _PINNED[shape] = torch.empty(shape, dtype=torch.float64).pin_memory()
    return _PINNED[shape]


def _dist():
    try:
        import torch.distributed as dist
    except ImportError:
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def _solve_shard(wind, props, cells, ndays, rad_dist, rad_res, prob_model, r_dur, r_number, r_dist, r_start, device,
                 wind_device_ptr, wind_shape, ids, sprd_factors=None, sprd_drift=(-25., 15.), projection=None, out_device=None):
    """This rank's proposals through ONE library call (``pkb_solve_batch``: kernel construction
    batched over groups of proposals, chains back to back on the device-resident kernels)."""
    import ctypes as C
    from . import _abi, _lib
    from . import ParasitoidModel as PM
    hp, dp, dl, mu_r, n_periods = unpack_proposal(props[0])
    a, keep = Run._solve_args(wind, ndays, hp, dp, dl, mu_r, n_periods, rad_dist, rad_res, prob_model, r_dur, r_number,
                              r_dist, r_start, False, False, True, wind_device_ptr, wind_shape,
                              sprd_factor=None if sprd_factors is None else float(sprd_factors[0]), sprd_drift=sprd_drift)
    if sprd_factors is not None:
        sf = np.ascontiguousarray(sprd_factors, dtype=np.float64)
        a.sprd_factors = _lib.dptr(sf)
        keep = (keep, sf)
    props = np.ascontiguousarray(props, dtype=np.float64)
    n = props.shape[0]
    status = np.zeros((n, ndays), dtype=np.int32)
    if out_device is not None:          # results straight into the caller's device tensor (the one it all-gathers)
        a.out_on_device = 1
        out, optr = out_device, C.cast(C.c_void_p(out_device.data_ptr()), _abi.c_double_p)
    else:
        out = np.empty((n, ndays, cells.shape[0]) if projection is None else (n, projection.nrows))
        optr = _lib.dptr(out)
    if projection is None:
        _lib.check(_lib.lib().pkb_solve_batch(_lib.ctx(device).h, C.byref(a), _lib.dptr(props), n, _lib.iptr(cells), cells.shape[0],
                                              optr, _lib.iptr(status)))
    else:
        # the likelihood projection of every proposal on the device (Bayes_Run.py:298-306): only its rows come back
        _lib.check(_lib.lib().pkb_solve_batch_projected(_lib.ctx(device).h, C.byref(a), _lib.dptr(props), n, _lib.iptr(cells),
                                                        cells.shape[0], C.byref(projection.c), optr, _lib.iptr(status)))
    del keep
    # the reference's assertion / warning sites, per (proposal, day) kernel (ParasitoidModel.py:529-599)
    for i in range(n):
        for d in np.nonzero(status[i])[0]:
            meta = _abi.DayMeta()
            meta.status = int(status[i, d])
            PM._raise_for_status(meta, int(d), ('proposal', int(ids[i])) + tuple(props[i]))
    return out


def solve_batch(wind, proposals, cells, ndays, rad_dist, rad_res, prob_model=False, r_dur=1, r_number=1.0,
                r_dist=None, r_start=None, device=None, group=None, wind_device_ptr=None, wind_shape=None,
                sprd_factor=None, sprd_drift=(-25., 15.), projection=None):
    """Solve every proposal and return the model at ``cells`` for all of them.

    wind:       (nd_wind, periods, 3) consecutive days (``Run.stack_wind``), shared by all proposals
    proposals:  (B, 15) array in ``PROPOSAL_FIELDS`` order
    cells:      (K, 2) int (row, col) sample cells of the domain
    sprd_factor: None, a scalar or (B,) -- the leading local-spread day of Bayes_Run.py:245-296 (its own sampled
                variable there), see ``Run.solve``
    projection: a ``Bayes_funcs.Projection`` (then ``cells`` is ignored, its own sample cells are used): every
                proposal's solution is folded into the values ``popdensity_to_emergence`` / ``popdensity_grid`` return,
                on the device; the result is (B, projection.nrows), ``projection.split(row)`` gives the arrays
    returns     (B, ndays, K) float64, identical on every rank (with NCCL: a view of a cached pinned buffer, overwritten
                by the next call with the same shape -- copy it to keep it)

    With an initialised ``torch.distributed`` process group the proposals are
    sharded over the ranks (``shard``) and the results all-gathered; without
    one this is a plain loop on one GPU."""
    proposals = np.asarray(proposals, dtype=float).reshape(-1, len(PROPOSAL_FIELDS))
    if projection is not None:
        if projection.ndays != ndays:
            raise ValueError('the projection was built for {} model days, the solve has {}'.format(projection.ndays, ndays))
        cells = projection.cells
    cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 2)
    B, K = proposals.shape[0], cells.shape[0]
    tail = (ndays, K) if projection is None else (projection.nrows,)
    dist = _dist()
    world = dist.get_world_size(group) if dist else 1
    rank = dist.get_rank(group) if dist else 0
    # the ranks finish together only if their shares cost about the same: deal the proposals out by estimated cost
    cost = proposal_cost(proposals) if world > 1 else None
    mine = shard(B, world, rank, cost)
    per = -(-B // world)                      # padded shard length
    sf = None if sprd_factor is None else np.broadcast_to(np.asarray(sprd_factor, dtype=float), (B,))
    shard_args = (wind, proposals[mine], cells, ndays, rad_dist, rad_res, prob_model, r_dur, r_number, r_dist, r_start, device,
                  wind_device_ptr, wind_shape, mine, None if sf is None else sf[mine], sprd_drift, projection)
    if world == 1:
        return _solve_shard(*shard_args)[:B] if mine else np.zeros((0,) + tail)
    import torch
    if dist.get_backend(group) == 'nccl':
        # the library writes this rank's results straight into the tensor that is all-gathered over NVLink; the only
        # host copy is the one of the gathered result
        dev = torch.device('cuda', device if device is not None else torch.cuda.current_device())
        t = torch.zeros((per,) + tail, dtype=torch.float64, device=dev)
        if mine:
            _solve_shard(*shard_args, out_device=t)
    else:
        local = np.zeros((per,) + tail)
        if mine:
            local[:len(mine)] = _solve_shard(*shard_args)
        t = torch.from_numpy(local)
    out = torch.empty((world * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)       # concatenated along dim 0 (gloo and nccl both accept this form)
    # back to proposal order where the data is (one gather on the device), then ONE copy into pinned host memory
    pos = np.empty(B, dtype=np.int64)
    for r in range(world):
        idx = shard(B, world, r, cost)
        pos[idx] = r * per + np.arange(len(idx))
    full = out.index_select(0, torch.from_numpy(pos).to(out.device))
    if full.is_cuda:
        host = _pinned(tuple(full.shape))
        host.copy_(full, non_blocking=True)
        torch.cuda.current_stream(full.device).synchronize()
        return host.numpy().copy() if _COPY_OUT else host.numpy()
    return full.numpy()
