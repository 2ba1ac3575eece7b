#!/bin/bash
set -o pipefail
timeout 900 python -m pytest tests/test_chain.py tests/test_fullsize.py tests/test_parity_full.py -m gpu -x -q -k "not c5 and not population" < /dev/null > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --no-extras < /dev/null > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; cut -c1-140 gpurun_out/r2l_bench.json
timeout 300 python bench.py --no-cpu-baseline --no-extras --opt tau_lag=2 < /dev/null > gpurun_out/r2l_bench_lag2.json 2> gpurun_out/r2l_bench_lag2.err; echo "bench lag2 rc=$?"; cut -c1-140 gpurun_out/r2l_bench_lag2.json
