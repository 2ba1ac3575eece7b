#!/bin/bash
# register cap 96 (tools/chainbench_r96: -DPKB_MAXNREG=96): more resident threads per SM
cb() { echo "== $*"; bin=$1; shift; env "$@" timeout 60 tools/chainbench_$bin 4097 361 10 < /dev/null;  }
cbk() { echo "== kalbar $*"; bin=$1; shift; env "$@" timeout 60 tools/chainbench_$bin 801 641 20 < /dev/null;  }
cb bfast A=1
cb r96 A=1
cb r96 PKB_FFT_T=192 PKB_COLS_T=192
cb r96 PKB_FFT_T=224 PKB_COLS_T=224
cb r96 PKB_FFT_T=224 PKB_COLS_T=192
cb r96 PKB_FFT_T=192 PKB_COLS_T=224
for t in 160 192 224; do timeout 60 tools/chainbench_r96 core 4704 $t 20 < /dev/null; done
timeout 60 tools/chainbench_bfast core 4704 160 20 < /dev/null
cbk bfast A=1
cbk r96 A=1
cbk r96 PKB_FFT_OCC=5
cbk r96 PKB_FFT_T=160 PKB_COLS_T=160
cbk r96 PKB_FFT_T=96 PKB_COLS_T=96 PKB_FFT_OCC=6
sum() {
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], 'days/s %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['roofline']['kernel'], '%.1f us' % (1000*d['roofline']['avg_launch_ms']), flush=True)
except Exception as e:
    print(sys.argv[2], 'FAILED', e, flush=True)
PY
}
c5() { tag=$1; shift; timeout 200 python bench.py --workload kalbar_batch512 --steps 1 --warmup 1 --no-cpu-baseline "$@" < /dev/null > gpurun_out/v_c5_$tag.json 2> gpurun_out/v_c5_$tag.err; sum gpurun_out/v_c5_$tag.json c5_$tag; }
c5 l4_st1 --opt batch_lanes=4
c5 l4_st0 --opt batch_lanes=4 --opt step_torus=0
c5 l6_st1 --opt batch_lanes=6
c5 l8_st1 --opt batch_lanes=8
