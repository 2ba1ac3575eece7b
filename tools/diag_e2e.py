"""Diagnostic (GPU): where the end-to-end time of the C4 solve goes -- device-resident solve against the same solve with
COO / CSR output pumped to the host, library phase times (pkb_timing) beside the wall clock."""
import sys, time, warnings, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from parasitoids_b200 import _lib, Run
wind, wind_data, days, rad_dist, rad_res = bench.load_workload('synthetic_4097x4097_60d')
ctx = _lib.ctx(0)
wpin = torch.from_numpy(wind).pin_memory().numpy()
warnings.simplefilter('ignore')
names = ['k_rows_fwd', 'k_cols', 'k_rows_inv', 'k_rows_fwd_win', 'k_cols_win', 'k_rows_inv_win', 'k_kernel_rows_win', 'k_emit_dense', 'k_row_scan', 'k_coo_write',
         'k_row_nnz', 'k_period', 'k_day_finalize']
for mode, prof in ((False, False), ('csr', False), ('coo', False), ('csr', True), (False, True)):
    for rep in range(3):
        if prof and rep == 2:
            ctx.profile_reset(); ctx.profile(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = Run.solve(wpin, len(days), bench.HPARAMS, bench.DPARAMS, bench.DLPARAMS, bench.MU_R, bench.N_PERIODS, rad_dist, rad_res,
                        want_coo=mode, device=0)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        tm = ctx.timing()
        nnz = None
        if mode == 'csr':
            nnz = int(res.csr_arrays()[0][-1]) if hasattr(res, 'csr_arrays') else None
        res.close()
    rec = {'want_coo': mode, 'profiled': prof, 'wall_ms': round((t1 - t0) * 1e3, 2), 'timing': {k: round(v, 2) for k, v in tm.items()}, 'nnz': nnz}
    if prof:
        ctx.profile(False)
        rec['kernel_ms'] = {n: [v[0], round(v[1], 2)] for n, v in ((n, ctx.profile_get(n)) for n in names) if v[0]}
    print(json.dumps(rec), flush=True)
