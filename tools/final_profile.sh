#!/bin/bash
# Round-end evidence run on one B200 (gpurun): GPU parity tests, smoke, the bench line, then -- only after
# the plain commands exited 0 -- the ncu launch list and one full capture of the three FFT kernels of a
# whole-torus chain step.  usage: tools/final_profile.sh <tag>   (outputs under gpurun_out/<tag>_*)
tag=${1:-r1u}
set -o pipefail
timeout 400 python -m pytest tests -m gpu -x -q < /dev/null > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke < /dev/null > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 500 python bench.py < /dev/null > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 < /dev/null > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "bench ref rc=$?"
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline < /dev/null > gpurun_out/${tag}_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 800 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline < /dev/null > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_cols|^k_rows_fwd|^k_rows_inv' -s 270 -c 3 -f -o gpurun_out/${tag}_prof \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline < /dev/null > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_*
