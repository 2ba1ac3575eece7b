import sys, warnings
sys.path.insert(0, '.')
import bench
from parasitoids_b200 import Run, _lib
wind, wind_data, days, rad_dist, rad_res = bench.load_workload('synthetic_4097x4097_60d')
ctx = _lib.ctx(0)
model = (bench.HPARAMS, bench.DPARAMS, bench.DLPARAMS, bench.MU_R, bench.N_PERIODS, rad_dist, rad_res)
warnings.simplefilter('ignore')
for i in range(3):
    ctx.profile_reset(); ctx.profile(True)
    try:
        r = Run.solve(wind, 60, *model, want_coo=False, keep_device=True); r.close()
    except Exception as e:
        print('solve failed (expected in experiments):', str(e)[:80])
    ctx.profile(False)
    print({k: round(ctx.profile_get(k)[1], 3) for k in ('k_period', 'k_day_finalize', 'k_drift', 'k_hprob')}, ctx.timing())
