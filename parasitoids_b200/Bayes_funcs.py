"""Drop-in for the reference's ``Bayes_funcs`` module: the likelihood projection of a
population-model solution, computed on the GPU.

``popdensity_to_emergence(modelsol, locinfo)`` and ``popdensity_grid(modelsol,
locinfo)`` keep the reference's signatures and return types (Bayes_funcs.py:20-180).
Both only read the model at a few cells (``locinfo.emerg_grids``, ``field_cells``,
``grid_cells``) and fold it with the incubation distribution, so the whole thing is a
small ordered linear map of the model at K sample cells.  ``Projection`` builds that
map from a ``LocInfo`` by walking the reference's own loops; the device evaluates it in
the reference's summation order (csrc/project.cuh), so the values are the reference's to
the last bit.  Three ways in:

* a list of scipy sparse matrices / arrays, as the reference is called (the sampled
  cells are uploaded);
* a ``Run.SolveResult`` whose dense days are still on the device (nothing but the
  projected values crosses PCIe);
* ``batch.solve_batch(..., projection=...)``: every proposal of a likelihood batch is
  projected on the device straight after its chain (Bayes_Run.py:298-306).

There is no CPU fallback; without the CUDA library these functions raise.
"""
import ctypes as C

import numpy as np

from . import _abi, _lib

# Bayes_funcs.py:10-18 -- oviposition-to-emergence delay, 19..25 days
incubation_time = np.array([0.05, 0.1, 0.2, 0.3, 0.2, 0.1, 0.05])
max_incubation_time = 25


def _obs_days(dframe):
    return np.array(dframe['datePR'].map(lambda t: t.days).unique())


class Projection(object):
    """The linear map from a model (``ndays`` daily grids) to everything ``popdensity_to_emergence`` and
    ``popdensity_grid`` return for ``locinfo``.

    ``cells``  (K, 2) int32 sample cells (row, col), each distinct cell once
    rows       one output value each, laid out as: release_emerg[0] (points x observation days, C order),
               release_emerg[1], ..., sentinel_emerg[0], ..., grid_counts (points x observation days)
    """

    def __init__(self, locinfo, ndays, with_emergence=True, with_grid=True):
        self.ndays = int(ndays)
        self._cell_id = {}
        self._sets = []            # list of lists of sample-cell indices
        self._rows = []            # list of groups; group = list of (day, set, weight)
        self.layout = []           # (kind, shape) per returned array, in row order
        if with_emergence:
            for nframe, dframe in enumerate(locinfo.release_DataFrames):                    # Bayes_funcs.py:30-88
                sets = [self._set([(r, c)]) for r, c in locinfo.emerg_grids[nframe]]
                self._emergence('release', sets, locinfo.collection_datesPR[nframe].days, _obs_days(dframe))
            for nframe, dframe in enumerate(locinfo.sent_DataFrames):                       # :91-143
                sets = [self._set([(int(r), int(c)) for r, c in locinfo.field_cells[f]]) for f in locinfo.sent_ids]
                self._emergence('sentinel', sets, locinfo.collection_datesPR[nframe].days, _obs_days(dframe))
        if with_grid:                                                                       # :156-180
            sets = [self._set([(int(r), int(c))]) for r, c in locinfo.grid_cells]
            nobs = len(locinfo.grid_obs_datesPR)
            for s in sets:
                for date in locinfo.grid_obs_datesPR:
                    self._check_day(date.days - 1)
                    self._rows.append([[(date.days - 1, s, 1.0)]])
            self.layout.append(('grid', (len(sets), nobs)))
        self.cells = np.array(sorted(self._cell_id, key=self._cell_id.get), dtype=np.int32).reshape(-1, 2)
        self._pack()

    # -- construction -----------------------------------------------------------------
    def _check_day(self, day):
        if not 0 <= day < self.ndays:
            raise IndexError('the projection needs model day {} but the solve has {} days'.format(day, self.ndays))

    def _set(self, cells):
        ids = []
        for rc in cells:
            rc = (int(rc[0]), int(rc[1]))
            if rc not in self._cell_id:
                self._cell_id[rc] = len(self._cell_id)
            ids.append(self._cell_id[rc])
        self._sets.append(ids)
        return len(self._sets) - 1

    def _emergence(self, kind, sets, collection_day, obs_datesPR):
        """Rows of one collection: for every point and observation date the emergence columns it sums
        (Bayes_funcs.py:79-85), each column accumulated over the feasible oviposition days (:58-71)."""
        start_day = max(collection_day - max_incubation_time, 0)
        # terms[e] = [(day, weight)] in the order the reference adds them to emerg_proj[:, e]
        terms = [[] for _ in range(max_incubation_time)]
        for day in range(start_day, collection_day):
            self._check_day(day)
            max_post_col = day + max_incubation_time - collection_day
            min_post_col = max(0, max_post_col + 1 - incubation_time.size)
            span_len = max_post_col - min_post_col + 1
            for e, w in zip(range(min_post_col, max_post_col + 1), incubation_time[-span_len:]):
                terms[e].append((day, float(w)))
        col_indices = obs_datesPR - collection_day
        bins = [range(0, int(col_indices[0]) + 1)]
        for n, col in enumerate(col_indices[1:]):
            bins.append(range(int(col_indices[n]) + 1, int(col) + 1))
        for s in sets:
            for cols in bins:
                cols = [e for e in cols if 0 <= e < max_incubation_time]       # numpy slicing clips
                self._rows.append([[(day, s, w) for day, w in terms[e]] for e in cols])
        self.layout.append((kind, (len(sets), len(bins))))

    def _pack(self):
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)         # noqa: E731
        self.set_ptr = i32(np.concatenate([[0], np.cumsum([len(s) for s in self._sets])]))
        self.set_cells = i32([c for s in self._sets for c in s])
        groups = [g for row in self._rows for g in row]
        self.row_ptr = i32(np.concatenate([[0], np.cumsum([len(row) for row in self._rows])]))
        self.grp_ptr = i32(np.concatenate([[0], np.cumsum([len(g) for g in groups])]))
        terms = [t for g in groups for t in g]
        self.term_day = i32([t[0] for t in terms])
        self.term_set = i32([t[1] for t in terms])
        self.term_w = np.ascontiguousarray([t[2] for t in terms], dtype=np.float64)
        self.nrows = len(self._rows)
        p = _abi.Projection()
        p.nsets, p.nrows, p.ngroups = len(self._sets), self.nrows, len(groups)
        p.set_ptr, p.set_cells = _lib.iptr(self.set_ptr), _lib.iptr(self.set_cells)
        p.row_ptr, p.grp_ptr = _lib.iptr(self.row_ptr), _lib.iptr(self.grp_ptr)
        p.term_day, p.term_set, p.term_w = _lib.iptr(self.term_day), _lib.iptr(self.term_set), _lib.dptr(self.term_w)
        self.c = p

    # -- evaluation -------------------------------------------------------------------
    def split(self, values):
        """One proposal's row values -> (release_emerg, sentinel_emerg, grid_counts) in the reference's shapes
        (grid_counts is None if the projection was built without it)."""
        values = np.asarray(values, dtype=float).ravel()
        rel, sen, grid, pos = [], [], None, 0
        for kind, shape in self.layout:
            n = shape[0] * shape[1]
            a = values[pos:pos + n].reshape(shape)
            pos += n
            if kind == 'release':
                rel.append(a)
            elif kind == 'sentinel':
                sen.append(a)
            else:
                grid = a
        return rel, sen, grid

    def sample(self, modelsol):
        """(ndays, K) model values at ``cells`` from a list of arrays / scipy sparse matrices."""
        out = np.empty((self.ndays, self.cells.shape[0]))
        r, c = self.cells[:, 0], self.cells[:, 1]
        for d in range(self.ndays):
            out[d] = np.asarray(modelsol[d][r, c]).ravel()
        return out

    def apply_samples(self, samples, device=None):
        """samples: (B, ndays, K) or (ndays, K) host array -> (B, nrows) / (nrows,)."""
        s = _lib.as_f64(samples)
        single = s.ndim == 2
        s = s.reshape((-1, self.ndays, self.cells.shape[0]))
        out = np.empty((s.shape[0], self.nrows))
        _lib.check(_lib.lib().pkb_project(_lib.ctx(device).h, C.byref(self.c), _lib.dptr(s), s.shape[0], self.ndays,
                                         self.cells.shape[0], _lib.dptr(out)))
        return out[0] if single else out

    def apply(self, modelsol, device=None):
        """Row values for one model: a ``Run.SolveResult`` with dense days on the device, or a list of daily grids."""
        from . import Run
        if isinstance(modelsol, Run.SolveResult):
            if modelsol.ndays < self.ndays:
                raise IndexError('the projection needs {} model days, the solve has {}'.format(self.ndays, modelsol.ndays))
            out = np.empty(self.nrows)
            _lib.check(_lib.lib().pkb_result_project(modelsol.h, _lib.iptr(self.cells), self.cells.shape[0], C.byref(self.c),
                                                    _lib.dptr(out)))
            return out
        return self.apply_samples(self.sample(modelsol), device)


def _ndays(modelsol):
    from . import Run
    return modelsol.ndays if isinstance(modelsol, Run.SolveResult) else len(modelsol)


def popdensity_to_emergence(modelsol, locinfo):
    """Expected emergence per collection, point / sentinel field and observation date (Bayes_funcs.py:20-153).
    Returns ``(release_emerg, sentinel_emerg)``: two lists with one (points x observation dates) array per collection."""
    proj = Projection(locinfo, _ndays(modelsol), with_grid=False)
    rel, sen, _ = proj.split(proj.apply(modelsol))
    return rel, sen


def popdensity_grid(modelsol, locinfo):
    """Model population at the release-field grid points on the observation days (Bayes_funcs.py:156-180)."""
    proj = Projection(locinfo, _ndays(modelsol), with_emergence=False)
    return proj.split(proj.apply(modelsol))[2]
