#!/bin/bash
set -o pipefail
timeout 1500 python -m pytest tests -m gpu -x -q < /dev/null > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest_gpu.log
timeout 600 python bench.py < /dev/null > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; cut -c1-140 gpurun_out/r2k_bench.json; tail -2 gpurun_out/r2k_bench.err
