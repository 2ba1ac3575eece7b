"""CPU oracle, likelihood projection: Bayes_funcs.py restated.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Plain numpy; pinned
against outputs of the reference module itself (tests/golden/bayes_funcs.npz,
oracle/make_golden.py:make_bayes_funcs).

``locinfo`` is any object with the attributes the reference reads
(Data_Import.LocInfo): collection_datesPR, emerg_grids, release_DataFrames,
sent_DataFrames, sent_ids, field_cells, grid_cells, grid_obs_datesPR.
"""
import numpy as np

# Bayes_funcs.py:10-18 -- oviposition-to-emergence delay, 19..25 days
incubation_time = np.array([0.05, 0.1, 0.2, 0.3, 0.2, 0.1, 0.05])
max_incubation_time = 25


def _obs_days(dframe):
    """Unique observation dates (days post release) in first-appearance order (Bayes_funcs.py:76)."""
    return np.array(dframe['datePR'].map(lambda t: t.days).unique())


def _project(point_values, collection_day, obs_datesPR):
    """Bayes_funcs.py:44-88 for one collection.  point_values(day) -> 1-D array over the points (grid cells or
    sentinel fields) of the model on that day."""
    start_day = max(collection_day - max_incubation_time, 0)                 # :44
    first = point_values(start_day) if start_day < collection_day else None
    npts = 0 if first is None else len(first)
    emerg_proj = np.zeros((npts, max_incubation_time))                       # :54-55
    for day in range(start_day, collection_day):                             # :58
        max_post_col = day + max_incubation_time - collection_day            # :62
        min_post_col = max(0, max_post_col + 1 - incubation_time.size)       # :63
        span_len = max_post_col - min_post_col + 1                           # :64
        vals = point_values(day)
        for n in range(npts):
            e_distrib = vals[n] * incubation_time                            # :70
            emerg_proj[n, min_post_col:max_post_col + 1] += e_distrib[-span_len:]     # :71
    col_indices = obs_datesPR - collection_day                               # :79
    out = np.zeros((npts, len(obs_datesPR)))
    out[:, 0] = emerg_proj[:, 0:col_indices[0] + 1].sum(axis=1)              # :82
    for n, col in enumerate(col_indices[1:]):                                # :83-85
        out[:, n + 1] = emerg_proj[:, col_indices[n] + 1:col + 1].sum(axis=1)
    return out


def popdensity_to_emergence(modelsol, locinfo):
    """Bayes_funcs.py:20-153.  modelsol: one 2-D array / sparse matrix per day."""
    release_emerg = []
    for nframe, dframe in enumerate(locinfo.release_DataFrames):
        cday = locinfo.collection_datesPR[nframe].days
        grid = locinfo.emerg_grids[nframe]
        release_emerg.append(_project(lambda day: np.array([modelsol[day][r, c] for r, c in grid]), cday, _obs_days(dframe)))
    sentinel_emerg = []
    for nframe, dframe in enumerate(locinfo.sent_DataFrames):
        cday = locinfo.collection_datesPR[nframe].days

        def fields(day):                                                     # :117-118
            return np.array([np.asarray(modelsol[day][locinfo.field_cells[f][:, 0], locinfo.field_cells[f][:, 1]]).sum()
                             for f in locinfo.sent_ids])
        sentinel_emerg.append(_project(fields, cday, _obs_days(dframe)))
    return release_emerg, sentinel_emerg


def popdensity_grid(modelsol, locinfo):
    """Bayes_funcs.py:156-180 -- model at the release-field grid points on the observation days (end of the day before)."""
    out = np.zeros((locinfo.grid_cells.shape[0], len(locinfo.grid_obs_datesPR)))
    for nday, date in enumerate(locinfo.grid_obs_datesPR):
        for n, (r, c) in enumerate(locinfo.grid_cells):
            out[n, nday] = modelsol[date.days - 1][r, c]
    return out
