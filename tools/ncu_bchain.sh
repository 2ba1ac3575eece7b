#!/bin/bash
# ncu --set full capture of one step of the batched chain (kb_rows_fwd, kb_cols, kb_rows_inv) in the middle of a 32-proposal Kalbar group
set -e
mkdir -p gpurun_out
PKB_BCHAIN_DEBUG=1 python tools/diag_bchain1.py 32 0 > gpurun_out/bchain_dbg.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:kb_ -s 40 -c 4 -o gpurun_out/bchain_full -f python tools/diag_bchain1.py 32 0 > gpurun_out/ncu_bchain.log 2>&1 || true
ncu -i gpurun_out/bchain_full.ncu-rep --page raw --csv > gpurun_out/bchain_full_raw.csv 2>/dev/null || true
tail -3 gpurun_out/ncu_bchain.log
grep bchain: gpurun_out/bchain_dbg.log | head -3
