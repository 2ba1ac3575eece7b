#!/bin/bash
# round 2, call E: all GPU tests, the full default bench (extras + CPU baseline), reference arm, ncu metrics of k_period
set -o pipefail
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 < /dev/null > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r2e_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke < /dev/null > gpurun_out/r2e_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2e_smoke.log
( time timeout 900 python bench.py < /dev/null > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2e_bench.json; tail -3 gpurun_out/r2e_bench.err
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 < /dev/null > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err ) 2>&1 | grep real; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/r2e_bench_ref.json; tail -3 gpurun_out/r2e_bench_ref.err
timeout 200 python tools/diag_phase1.py < /dev/null > gpurun_out/r2e_phase1_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
   --clock-control none -k regex:k_period -c 1 --csv --log-file gpurun_out/r2e_ncu_phase1.csv python tools/diag_phase1.py < /dev/null > gpurun_out/r2e_ncu_phase1.log 2>&1; echo "ncu rc=$?"; tail -8 gpurun_out/r2e_ncu_phase1.csv
