"""Diagnostic (GPU): the 801^2 presets (C1, C2, C3) through Run.solve with CSR output -- wall clock, library phase times and
per-kernel device times."""
import sys, time, warnings, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from parasitoids_b200 import _lib, Run
ctx = _lib.ctx(0)
warnings.simplefilter('ignore')
names = ['k_rows_fwd', 'k_cols', 'k_rows_inv', 'k_rows_fwd_win', 'k_cols_win', 'k_rows_inv_win', 'k_kernel_rows_win', 'k_kernel_rows', 'k_kernel_rows_batch',
         'k_emit_dense', 'k_emit_population', 'k_row_scan', 'k_coo_write', 'k_row_nnz', 'k_period', 'k_day_finalize', 'k_drift', 'k_hprob', 'k_place_kernel',
         'k_set_ctrl', 'k_zero_outside', 'k_copy_domain']
for name, site, model in (('c1', 'kalbar', 'prob'), ('c2', 'kalbar', 'pop'), ('c3_prob', 'carnarvon', 'prob'), ('c3_pop', 'carnarvon', 'pop')):
    wind, wind_data, days, rd, rr = bench.site_wind(site)
    r_dur, r_number, r_start = (1, 130000.0, None) if site == 'kalbar' else (5, 40000.0, 0.354)
    skw = dict(prob_model=True) if model == 'prob' else dict(prob_model=False, r_dur=r_dur, r_number=r_number, r_dist=[1.0 / r_dur] * r_dur, r_start=r_start)
    wp = torch.from_numpy(wind).pin_memory().numpy()
    m = (bench.HPARAMS, bench.DPARAMS, bench.DLPARAMS, bench.MU_R, bench.N_PERIODS, rd, rr)
    for prof in (False, True):
        for rep in range(3):
            if prof and rep == 2:
                ctx.profile_reset(); ctx.profile(True)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = Run.solve(wp, len(days), *m, want_coo='csr', device=0, **skw)
            res.csr_arrays()
            torch.cuda.synchronize(); t1 = time.perf_counter()
            tm = ctx.timing()
            info = dict(P=res.P, N=res.N, window_steps=res.window_steps(), flags=int(np.sum(res.flags())))
            res.close()
        rec = {'config': name, 'days': len(days), 'profiled': prof, 'wall_ms': round((t1 - t0) * 1e3, 2), 'timing': {k: round(v, 2) for k, v in tm.items()}, **info}
        if prof:
            ctx.profile(False)
            rec['kernel_ms'] = {n: [v[0], round(v[1], 2)] for n, v in ((n, ctx.profile_get(n)) for n in names) if v[0]}
        print(json.dumps(rec), flush=True)
