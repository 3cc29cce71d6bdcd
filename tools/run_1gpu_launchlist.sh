#!/bin/bash
mkdir -p gpurun_out
timeout 100 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r2z_ncu_launches_bench.csv python bench.py --steps 10 --warmup 3 --pcg-iters 5 --no-cpu --no-e2e --no-trilinear \
  > gpurun_out/r2z_bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2z_ncu_launches_bench.csv > gpurun_out/r2z_ncu_launches_bench_summary.txt 2>&1
cat gpurun_out/r2z_ncu_launches_bench_summary.txt
tail -2 gpurun_out/r2z_bench_under_ncu.log | cut -c1-200
