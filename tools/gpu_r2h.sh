#!/bin/bash
set -o pipefail
timeout 300 python bench.py --single-solve --steps 5 --warmup 2 < /dev/null > gpurun_out/r2h_single_n1.json 2> gpurun_out/r2h_single_n1.err; echo "single n1 rc=$?"; cut -c1-250 gpurun_out/r2h_single_n1.json; tail -2 gpurun_out/r2h_single_n1.err
timeout 600 python -m pytest tests/test_chain.py tests/test_multi.py -m gpu -x -q < /dev/null > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log
