"""Diagnostic (GPU): where the time of the 512-proposal Kalbar likelihood batch goes.
lanes = 1 with per-kernel events: kernel-time sum against wall (= launch gaps + host);
lanes = 4: the production setting."""
import sys, time, warnings, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from parasitoids_b200 import _lib, batch
nprop = int(sys.argv[1]) if len(sys.argv) > 1 else 128
wind, wind_data, days, rad_dist, rad_res = bench.load_workload('kalbar_batch512')
ctx = _lib.ctx(0)
wd = torch.from_numpy(wind).cuda(0)
cells = np.random.default_rng(7).integers(0, 2 * rad_res + 1, (1024, 2)).astype(np.int32)
props = bench.prior_proposals(512)[:nprop]
warnings.simplefilter('ignore')
kw = dict(prob_model=False, r_dur=1, r_number=130000.0, device=0, wind_device_ptr=wd.data_ptr(), wind_shape=wind.shape)
names = ['k_rows_fwd', 'k_cols', 'k_rows_inv', 'k_rows_fwd_win', 'k_cols_win', 'k_rows_inv_win', 'k_kernel_rows_win', 'k_kernel_rows',
         'k_kernel_rows_batch', 'k_emit_population_cells', 'k_emit_dense_cells', 'k_copy_domain_cells', 'k_step_finalize', 'k_zero_pad', 'k_period',
         'k_day_finalize', 'k_drift', 'k_hprob', 'k_bvn_setup', 'k_place_kernel', 'k_set_ctrl', 'k_stencil', 'k_row_stats']
for lanes, prof, thr in ((1, True, 0), (1, False, 0), (4, False, 0), (4, False, 0), (4, False, 1), (4, False, 1), (6, False, 1), (8, False, 1)):
    ctx.set_option('batch_lanes', lanes)
    ctx.set_option('batch_threads', thr)
    batch.solve_batch(None, props[:8], cells, 18, rad_dist, rad_res, **kw)
    if prof:
        ctx.profile_reset(); ctx.profile(True)
    l0 = ctx.launch_count()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    batch.solve_batch(None, props, cells, 18, rad_dist, rad_res, **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rec = {'lanes': lanes, 'threads': thr, 'profiled': prof, 'nprop': nprop, 'wall_ms': round((t1 - t0) * 1e3, 1), 'launches': ctx.launch_count() - l0,
           'days_per_s': round(nprop * 18 / (t1 - t0), 1)}
    if prof:
        ctx.profile(False)
        k = {n: ctx.profile_get(n) for n in names}
        rec['kernel_ms'] = {n: [v[0], round(v[1], 2)] for n, v in k.items() if v[0]}
        rec['kernel_ms_sum'] = round(sum(v[1] for v in k.values()), 1)
    print(json.dumps(rec))
