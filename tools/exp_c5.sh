#!/bin/bash
# C5 (kalbar_batch512) tuning sweep: CTA size / resident-CTA cap of the FFT kernels at Kalbar's torus sizes
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --workload kalbar_batch512 --steps 1 --warmup 1 --no-cpu-baseline < /dev/null > gpurun_out/c5_$tag.json 2> gpurun_out/c5_$tag.err
  python - gpurun_out/c5_$tag.json $tag <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    km=d.get('roofline_chain') or {}
    print(sys.argv[2], 'days/s %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['roofline']['kernel'], '%.1f us' % (1000*d['roofline']['avg_launch_ms']), flush=True)
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run base A=1
run t64o8 PKB_FFT_T=64 PKB_COLS_T=64 PKB_FFT_OCC=8
run t96o5 PKB_FFT_T=96 PKB_COLS_T=96 PKB_FFT_OCC=5
run t128o4c64 PKB_COLS_T=64 PKB_FFT_OCC=8
run t256 PKB_FFT_T=256 PKB_COLS_T=256
