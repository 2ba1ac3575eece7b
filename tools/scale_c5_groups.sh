#!/bin/bash
# C5 on 8 GPUs: kernel-construction group size of pkb_solve_batch (64 proposals per rank)
for g in 16 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$((g%10)) bench.py --gpus 8 --steps 3 --warmup 2 --workload kalbar_batch512 --opt batch_group=$g < /dev/null > gpurun_out/c5n8_g$g.json 2> gpurun_out/c5n8_g$g.err
python - gpurun_out/c5n8_g$g.json $g <<'PY'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][-1]
    print('group', sys.argv[2], 'days/s %.0f e2e %.0f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
except Exception as e: print('group', sys.argv[2], 'FAILED', e)
PY
done
