#!/bin/bash
set -o pipefail
timeout 900 python -m pytest tests/test_chain.py tests/test_fullsize.py tests/test_parity_full.py -m gpu -x -q -k "not c5 and not population" < /dev/null > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2j_pytest_gpu.log
for opt in "tau_lag=2" "tau_lag=1" "tau_lag=3" "tau_windows=0"; do
  timeout 300 python bench.py --no-cpu-baseline --no-extras --opt $opt < /dev/null > gpurun_out/r2j_bench_$opt.json 2> gpurun_out/r2j_bench_$opt.err; echo "bench $opt rc=$?"; cut -c1-140 gpurun_out/r2j_bench_$opt.json
done
python - <<'PY'
import sys, warnings
sys.path.insert(0,'.')
import bench
from parasitoids_b200 import Run
wind, wd, days, rd, rr = bench.load_workload('synthetic_4097x4097_60d')
warnings.simplefilter('ignore')
r = Run.solve(wind, 60, bench.HPARAMS, bench.DPARAMS, bench.DLPARAMS, bench.MU_R, 30, rd, rr, want_coo=False, keep_device=True)
print('window steps', r.window_steps(), 'spectral', r.spectral_steps())
for d in range(0, 60, 3): print(d, r.regions()[d])
PY
