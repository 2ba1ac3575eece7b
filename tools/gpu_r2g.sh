#!/bin/bash
# round 2, call G (2 GPUs): single-solve strong scaling N = 1, 2; the standard bench at N = 2 (extras: c5 strong scaling)
set -o pipefail
timeout 300 python bench.py --single-solve --steps 5 --warmup 2 < /dev/null > gpurun_out/r2g_single_n1.json 2> gpurun_out/r2g_single_n1.err; echo "single n1 rc=$?"; cut -c1-250 gpurun_out/r2g_single_n1.json; tail -2 gpurun_out/r2g_single_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --single-solve --steps 5 --warmup 2 < /dev/null > gpurun_out/r2g_single_n2.json 2> gpurun_out/r2g_single_n2.err; echo "single n2 rc=$?"; cut -c1-250 gpurun_out/r2g_single_n2.json; tail -3 gpurun_out/r2g_single_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 3 --warmup 3 < /dev/null > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-250 gpurun_out/r2g_bench_n2.json; tail -3 gpurun_out/r2g_bench_n2.err
