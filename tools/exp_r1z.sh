#!/bin/bash
sum() {
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][-1]
    rc=d.get('roofline_chain') or {}
    print(sys.argv[2], 'days/s %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['roofline']['kernel'], '%.1f us' % (1000*d['roofline']['avg_launch_ms']),
          'chain_ms/day %.4f' % rc.get('chain_ms_per_day', 0), {k: round(v, 2) for k, v in (rc.get('kernel_ms') or {}).items() if v > 0.3}, flush=True)
except Exception as e:
    print(sys.argv[2], 'FAILED', e, flush=True)
PY
}
b() { tag=$1; shift; timeout 250 python bench.py --no-cpu-baseline "$@" < /dev/null > gpurun_out/z_$tag.json 2> gpurun_out/z_$tag.err; sum gpurun_out/z_$tag.json $tag; }
timeout 400 python -m pytest tests -m gpu -x -q < /dev/null > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
b c4_desc1
b c4_desc0 --opt rows_desc=0
b c4_desc1b
b c4_desc0b --opt rows_desc=0
b c4_desc1_nost --opt step_torus=0
b c4_desc0_nost --opt step_torus=0 --opt rows_desc=0
