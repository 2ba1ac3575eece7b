#!/bin/bash
# multi-GPU evidence: the standard bench line at N GPUs (extras: c5 strong scaling, weak_batch); with a second argument also
# the sharded single solve.  usage: tools/multi_gpu_runs.sh N [tag] [single]
N=${1:-8}; tag=${2:-r2r}
set -o pipefail
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 3 --warmup 3 < /dev/null > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench n$N rc=$?"; cut -c1-200 gpurun_out/${tag}_bench_n$N.json; tail -2 gpurun_out/${tag}_bench_n$N.err
if [ -n "$3" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --single-solve --steps 5 --warmup 2 < /dev/null > gpurun_out/${tag}_single_n$N.json 2> gpurun_out/${tag}_single_n$N.err; echo "single n$N rc=$?"; cut -c1-200 gpurun_out/${tag}_single_n$N.json; tail -2 gpurun_out/${tag}_single_n$N.err
fi
