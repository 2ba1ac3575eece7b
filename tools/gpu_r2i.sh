#!/bin/bash
set -o pipefail
timeout 900 python -m pytest tests/test_chain.py tests/test_fullsize.py tests/test_parity_full.py tests/test_multi.py -m gpu -x -q -k "not c5 and not population" < /dev/null > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2i_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --no-extras < /dev/null > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2i_bench.json
timeout 300 python bench.py --no-cpu-baseline --no-extras --opt spectral_rows=0 < /dev/null > gpurun_out/r2i_bench_norows.json 2> gpurun_out/r2i_bench_norows.err; echo "bench norows rc=$?"; cut -c1-200 gpurun_out/r2i_bench_norows.json
python - <<'PY'
import sys, warnings
sys.path.insert(0,'.')
import bench
from parasitoids_b200 import Run
wind, wd, days, rd, rr = bench.load_workload('synthetic_4097x4097_60d')
warnings.simplefilter('ignore')
r = Run.solve(wind, 60, bench.HPARAMS, bench.DPARAMS, bench.DLPARAMS, bench.MU_R, 30, rd, rr, want_coo=False, keep_device=True)
print('row windows', r.row_windows())
print('spectral', r.spectral_steps())
print('padabs', ['%.0e' % r.day_meta(d)[1].padabs for d in range(60)])
PY
