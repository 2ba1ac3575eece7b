"""Diagnostic (GPU): ONE group of the Kalbar likelihood batch (no pipelining: kernel construction, then the chains), per-kernel
device times of the batched chain path and of the per-proposal path."""
import sys, time, warnings, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from parasitoids_b200 import _lib, batch
_lib.LIB_PATH = os.environ.get('PKB_DIAG_LIB', _lib.LIB_PATH)      # (tuning builds of the library)
nprop = int(sys.argv[1]) if len(sys.argv) > 1 else 32
off = int(sys.argv[2]) if len(sys.argv) > 2 else 0
wind, wind_data, days, rad_dist, rad_res = bench.site_wind('kalbar')
ctx = _lib.ctx(0)
wd = torch.from_numpy(wind).cuda(0)
cells = np.random.default_rng(7).integers(0, 2 * rad_res + 1, (1024, 2)).astype(np.int32)
props = bench.prior_proposals(512)[off:off + nprop]
warnings.simplefilter('ignore')
kw = dict(prob_model=False, r_dur=1, r_number=130000.0, device=0, wind_device_ptr=wd.data_ptr(), wind_shape=wind.shape)
nd = len(days)
names = ['kb_rows_fwd', 'kb_cols', 'kb_rows_inv', 'kb_finish', 'kb_init', 'k_rows_fwd', 'k_cols', 'k_rows_inv', 'k_rows_fwd_win', 'k_cols_win',
         'k_rows_inv_win', 'k_kernel_rows_win', 'k_kernel_rows', 'k_kernel_rows_batch', 'k_emit_population_cells', 'k_period', 'k_day_finalize',
         'k_drift', 'k_hprob', 'k_bvn_setup', 'k_place_kernel', 'k_set_ctrl', 'k_stencil', 'k_row_stats', 'k_step_finalize']
ctx.set_option('batch_group', max(nprop, 1))
for chain, prof in ((1, False), (1, False), (1, True), (0, False), (0, True)):
    ctx.set_option('batch_chain', chain)
    if prof:
        ctx.profile_reset(); ctx.profile(True)
    l0 = ctx.launch_count()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = np.array(batch.solve_batch(None, props, cells, nd, rad_dist, rad_res, **kw))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rec = {'batch_chain': chain, 'profiled': prof, 'nprop': nprop, 'wall_ms': round((t1 - t0) * 1e3, 1),
           'launches': ctx.launch_count() - l0, 'days_per_s': round(nprop * nd / (t1 - t0), 1), 'timing_ms': {k: round(v, 2) for k, v in ctx.timing().items()}}
    if prof:
        ctx.profile(False)
        k = {n: ctx.profile_get(n) for n in names}
        rec['kernel_ms'] = {n: [v[0], round(v[1], 2)] for n, v in k.items() if v[0]}
        rec['kernel_ms_sum'] = round(sum(v[1] for v in k.values()), 1)
    print(json.dumps(rec), flush=True)
