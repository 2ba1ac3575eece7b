#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q < /dev/null > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 100 python __graft_entry__.py smoke < /dev/null 2>&1 | tail -1
timeout 500 python bench.py < /dev/null > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 250 python bench.py --no-cpu-baseline < /dev/null > gpurun_out/final_bench2.json 2>> gpurun_out/final_bench.err
timeout 200 python bench.py --workload kalbar_batch512 --steps 2 --warmup 1 --no-cpu-baseline < /dev/null > gpurun_out/final_bench_c5.json 2>> gpurun_out/final_bench.err
python - <<'PY'
import json
for f in ('final_bench','final_bench2','final_bench_c5'):
    try:
        d=[json.loads(l) for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1]
        rc=d.get('roofline_chain') or {}
        print(f, 'days/s %.0f e2e %.0f launches %d' % (d['value'], d['e2e']['value'], d['gpu_launches']), d['roofline']['kernel'], '%.3f' % d['roofline']['frac'], rc.get('frac'), {k: round(v, 2) for k, v in (rc.get('kernel_ms') or {}).items() if v > 0.3})
    except Exception as e: print(f, 'FAILED', e)
PY
