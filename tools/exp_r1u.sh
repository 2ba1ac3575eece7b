#!/bin/bash
# round-1 late experiments: GPU parity tests, C4 with/without the per-step torus, C5 lanes / CTA geometry
sum() {
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    rc=d.get('roofline_chain') or {}
    print(sys.argv[2], 'days/s %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['roofline']['kernel'], '%.1f us' % (1000*d['roofline']['avg_launch_ms']),
          'chain_ms/day %.4f' % rc.get('chain_ms_per_day', 0), {k: round(v, 2) for k, v in (rc.get('kernel_ms') or {}).items() if v > 0.3}, flush=True)
except Exception as e:
    print(sys.argv[2], 'FAILED', e, flush=True)
PY
}
c4() { tag=$1; shift; timeout 200 python bench.py --no-cpu-baseline "$@" < /dev/null > gpurun_out/u_c4_$tag.json 2> gpurun_out/u_c4_$tag.err; sum gpurun_out/u_c4_$tag.json c4_$tag; }
c5() { tag=$1; shift; env "$@" timeout 200 python bench.py --workload kalbar_batch512 --steps 1 --warmup 1 --no-cpu-baseline $C5OPT < /dev/null > gpurun_out/u_c5_$tag.json 2> gpurun_out/u_c5_$tag.err; sum gpurun_out/u_c5_$tag.json c5_$tag; }
cb() { echo "== $*"; for k in 361 241; do env $2 timeout 60 tools/chainbench_$1 4097 $k 10 < /dev/null; done; env $2 timeout 60 tools/chainbench_$1 801 641 20 < /dev/null; }
cb nobfast PKB_NO_SLOTS=1
cb bfast PKB_NO_SLOTS=1
cb bfast A=1
timeout 300 python -m pytest tests -m gpu -x -q < /dev/null > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
c4 default
c4 nosteptorus --opt step_torus=0
C5OPT="--opt batch_lanes=1" c5 lanes1 A=1
C5OPT="" c5 lanes2 A=1
C5OPT="--opt batch_lanes=3" c5 lanes3 A=1
C5OPT="--opt batch_lanes=4" c5 lanes4 A=1
C5OPT="" c5 t64o8 PKB_FFT_T=64 PKB_COLS_T=64 PKB_FFT_OCC=8
C5OPT="" c5 t96o5 PKB_FFT_T=96 PKB_COLS_T=96 PKB_FFT_OCC=5
C5OPT="" c5 t256 PKB_FFT_T=256 PKB_COLS_T=256
