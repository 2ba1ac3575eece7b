#!/bin/bash
# ncu --set full capture of the kernel-construction kernels of one 32-proposal Kalbar group (k_period, k_day_finalize)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_period|k_day_finalize" -s 2 -c 2 -o gpurun_out/p1c5_full -f python tools/diag_bchain1.py 32 0 > gpurun_out/ncu_p1c5.log 2>&1
ncu -i gpurun_out/p1c5_full.ncu-rep --page raw --csv > gpurun_out/p1c5_full_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_p1c5.log
