#!/bin/bash
# round 2, call B: spectral-resident steps -- parity, bench with and without, ncu capture
set -o pipefail
timeout 900 python -m pytest tests/test_chain.py tests/test_fullsize.py tests/test_parity_full.py -m gpu -x -q -k "not c5 and not population" < /dev/null > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline < /dev/null > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2b_bench.json
timeout 300 python bench.py --no-cpu-baseline --opt spectral=0 < /dev/null > gpurun_out/r2b_bench_nospec.json 2> gpurun_out/r2b_bench_nospec.err; echo "bench nospec rc=$?"; cut -c1-300 gpurun_out/r2b_bench_nospec.json
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline < /dev/null > gpurun_out/r2b_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_cols|^k_rows_fwd|^k_rows_inv' -s 270 -c 3 -f -o gpurun_out/r2b_prof \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline < /dev/null > gpurun_out/r2b_ncu_full.log 2>&1; echo "ncu full rc=$?"
