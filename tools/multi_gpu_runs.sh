#!/bin/bash
# multi-GPU evidence: the standard bench line at N GPUs (extras: c5 strong scaling, weak_batch) and the sharded single solve
N=${1:-8}
set -o pipefail
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 3 --warmup 3 < /dev/null > gpurun_out/r2o_bench_n$N.json 2> gpurun_out/r2o_bench_n$N.err; echo "bench n$N rc=$?"; cut -c1-200 gpurun_out/r2o_bench_n$N.json; tail -2 gpurun_out/r2o_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --single-solve --steps 5 --warmup 2 < /dev/null > gpurun_out/r2o_single_n$N.json 2> gpurun_out/r2o_single_n$N.err; echo "single n$N rc=$?"; cut -c1-200 gpurun_out/r2o_single_n$N.json; tail -2 gpurun_out/r2o_single_n$N.err
