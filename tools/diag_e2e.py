import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from parasitoids_b200 import Run, _lib
wind, wind_data, days, rad_dist, rad_res = bench.load_workload('synthetic_4097x4097_60d')
ctx = _lib.ctx(0)
model = (bench.HPARAMS, bench.DPARAMS, bench.DLPARAMS, bench.MU_R, bench.N_PERIODS, rad_dist, rad_res)
wp = torch.from_numpy(wind).pin_memory().numpy()
for i in range(5):
    t0 = time.perf_counter()
    r = Run.solve(wp, 60, *model, want_coo=True)
    t1 = time.perf_counter()
    print('coo  wall %.1f ms' % ((t1 - t0) * 1e3), {k: round(v, 2) for k, v in ctx.timing().items()})
    r.close()
for i in range(3):
    t0 = time.perf_counter()
    r = Run.solve(wp, 60, *model, want_coo=False, keep_device=True)
    t1 = time.perf_counter()
    print('dev  wall %.1f ms' % ((t1 - t0) * 1e3), {k: round(v, 2) for k, v in ctx.timing().items()})
    r.close()
