#!/bin/bash
# round 2, call A: GPU parity tests (incl. the new full-size ones), bench baseline of the day, cuFFT yardstick, chainbench baseline
set -o pipefail
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt; free -g >> gpurun_out/r2a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 < /dev/null > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2a_pytest_gpu.log
timeout 120 tools/cufft_yardstick > gpurun_out/r2a_cufft.jsonl 2> gpurun_out/r2a_cufft.err; echo "cufft rc=$?"; cat gpurun_out/r2a_cufft.jsonl
timeout 200 tools/chainbench_base 4097 361 10 > gpurun_out/r2a_chainbench.log 2>&1; echo "chainbench rc=$?"; cat gpurun_out/r2a_chainbench.log
timeout 100 tools/chainbench_base core 4704 160 20 >> gpurun_out/r2a_chainbench.log 2>&1
timeout 600 python bench.py < /dev/null > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2a_bench.json
timeout 400 python bench.py --workload kalbar_batch512 --no-cpu-baseline --steps 3 --warmup 1 < /dev/null > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err; echo "bench c5 rc=$?"; cut -c1-400 gpurun_out/r2a_bench_c5.json
