#!/bin/bash
bash tools/final_profile.sh r1u
for l in 4; do
timeout 200 python bench.py --workload kalbar_batch512 --steps 2 --warmup 1 --no-cpu-baseline < /dev/null > gpurun_out/r1u_bench_c5_kalbar_batch512.json 2> gpurun_out/r1u_bench_c5.err; echo "c5 rc=$?"
done
python - <<'PY'
import json
for f in ('gpurun_out/r1u_bench.json','gpurun_out/r1u_bench_c5_kalbar_batch512.json'):
    try:
        d=json.load(open(f)); print(f, 'days/s %.0f e2e %.0f launches %d' % (d['value'], d['e2e']['value'], d['gpu_launches']), d['roofline']['kernel'], '%.3f' % d['roofline']['frac'], (d.get('roofline_chain') or {}).get('frac'))
    except Exception as e: print(f, 'FAILED', e)
PY
