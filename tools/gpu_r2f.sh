#!/bin/bash
# round 2, call F (1 GPU): new tests on the GPU, single-solve path at N = 1, parity of the exact accumulation
set -o pipefail
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 < /dev/null > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2f_pytest_gpu.log
timeout 300 python bench.py --single-solve --steps 5 --warmup 2 < /dev/null > gpurun_out/r2f_single_n1.json 2> gpurun_out/r2f_single_n1.err; echo "single n1 rc=$?"; cat gpurun_out/r2f_single_n1.json | cut -c1-900; tail -3 gpurun_out/r2f_single_n1.err
timeout 300 python bench.py --no-cpu-baseline --no-extras < /dev/null > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2f_bench.json
python tools/diag_c5.py 128 > gpurun_out/r2f_diag_c5.log 2>&1; tail -4 gpurun_out/r2f_diag_c5.log | cut -c1-400
