#!/bin/bash
# Evidence run on one B200 (gpurun): GPU parity tests, smoke, the bench line (with extras and the CPU baseline), the
# reference arm, then -- only after the plain command exited 0 -- the ncu launch list and one full capture of the three
# FFT kernels of one chain step.  usage: tools/final_profile.sh <tag>   (outputs under gpurun_out/<tag>_*)
tag=${1:-r2}
set -o pipefail
timeout 1500 python -m pytest tests -m gpu -x -q < /dev/null > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke < /dev/null > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${tag}_clocks.csv &
smi=$!
timeout 900 python bench.py < /dev/null > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
kill $smi
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 < /dev/null > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "bench ref rc=$?"
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras < /dev/null > gpurun_out/${tag}_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 900 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras < /dev/null > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_cols|^k_rows_fwd|^k_rows_inv' -s 270 -c 3 -f -o gpurun_out/${tag}_prof \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras < /dev/null > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_*
