#!/bin/bash
set -o pipefail
python tools/diag_spec.py > gpurun_out/r2c_diag_spec.log 2>&1; cat gpurun_out/r2c_diag_spec.log
timeout 900 python -m pytest tests/test_phase1.py tests/test_chain.py tests/test_batch.py -m gpu -q < /dev/null > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2c_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline < /dev/null > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2c_bench.json
timeout 400 python bench.py --workload kalbar_batch512 --no-cpu-baseline --steps 3 --warmup 1 < /dev/null > gpurun_out/r2c_bench_c5.json 2> gpurun_out/r2c_bench_c5.err; echo "bench c5 rc=$?"; cut -c1-200 gpurun_out/r2c_bench_c5.json
