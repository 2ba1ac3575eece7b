import sys, time, warnings
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from parasitoids_b200 import Run, _lib, batch
wind, wind_data, days, rad_dist, rad_res = bench.load_workload('synthetic_4097x4097_60d')
ctx = _lib.ctx(0)
model = (bench.HPARAMS, bench.DPARAMS, bench.DLPARAMS, bench.MU_R, bench.N_PERIODS, rad_dist, rad_res)
wd = torch.from_numpy(wind).cuda(0)
cells = np.random.default_rng(7).integers(0, 4097, (1024, 2)).astype(np.int32)
H = bench.HPARAMS
prop = np.array([[H[1], H[2], H[3], H[4], H[5], H[6], *bench.DPARAMS, *bench.DLPARAMS, H[0], bench.N_PERIODS, bench.MU_R]])
warnings.simplefilter('ignore')
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = Run.solve(None, 60, *model, want_coo=False, keep_device=True, wind_device_ptr=wd.data_ptr(), wind_shape=wind.shape); r.close()
    t1 = time.perf_counter()
    out = batch.solve_batch(None, prop, cells, 60, rad_dist, rad_res, prob_model=True, device=0, wind_device_ptr=wd.data_ptr(), wind_shape=wind.shape)
    t2 = time.perf_counter()
    print('solve %.2f ms   solve_batch(B=1) %.2f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), ctx.timing())
