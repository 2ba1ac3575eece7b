"""Diagnostic (GPU): direct stencil against the FFT path for small kernel radii -- the measurement behind the
`stencil_max_radius` default (BASELINE.json north_star item 2).  Per radius m and domain side D: device time of one chain step
(CUDA events around pkb_chain_conv's kernels via the library's per-kernel profile), both paths, and the HBM GB/s each reaches on
the bytes it has to move at least (read the P^2 state once, write it once)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from scipy import sparse
from parasitoids_b200 import cuda_lib, _lib
ctx = _lib.ctx(0)
rng = np.random.default_rng(0)
names = ['k_stencil', 'k_row_stats', 'k_step_finalize', 'k_kernel_rows', 'k_rows_fwd', 'k_cols', 'k_rows_inv']
for D in (801, 4097):
    A = rng.random((D, D)); A /= A.sum()
    for m in (1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24):
        B = rng.random((2 * m + 1, 2 * m + 1)); B /= B.sum()
        rec = {'D': D, 'm': m}
        for path, smax in (('stencil', 24), ('fft', -1)):
            ctx.set_option('stencil_max_radius', smax)
            s = cuda_lib.CudaSolve(sparse.coo_matrix(A), [2 * 24 + 1, 2 * 24 + 1])
            for _ in range(2):
                s.fftconv2(sparse.csr_matrix(B))
            ctx.profile_reset(); ctx.profile(True)
            reps = 5
            for _ in range(reps):
                s.fftconv2(sparse.csr_matrix(B))
            ctx.profile(False)
            us = sum(ctx.profile_get(k)[1] for k in names) * 1000.0 / reps
            P = s.pad_shape[0]
            rec[path + '_us'] = round(us, 1)
            rec[path + '_gbs'] = round(16.0 * P * P / us / 1e3, 1)
            s.close()
        print(json.dumps(rec))
ctx.set_option('stencil_max_radius', 3)
