#!/bin/bash
# usage: tools/scale_run.sh N [workload]   -- bench.py on N GPUs of one box the way the driver launches it
N=$1; W=${2:-synthetic_4097x4097_60d}
if [ "$N" = 1 ]; then
  timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --workload $W --no-cpu-baseline < /dev/null > gpurun_out/scale_${W}_n$N.json 2> gpurun_out/scale_${W}_n$N.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --workload $W < /dev/null > gpurun_out/scale_${W}_n$N.json 2> gpurun_out/scale_${W}_n$N.err
fi
echo "rc=$?"; tail -c 1500 gpurun_out/scale_${W}_n$N.json; tail -5 gpurun_out/scale_${W}_n$N.err
