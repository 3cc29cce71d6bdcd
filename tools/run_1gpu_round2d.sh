#!/bin/bash
mkdir -p gpurun_out
LIBP_OGS_TIMING=1 timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2t_bench_1gpu.json 2> gpurun_out/r2t_bench_1gpu.err
grep "ogs setup" gpurun_out/r2t_bench_1gpu.err | head -40
cut -c1-900 gpurun_out/r2t_bench_1gpu.json
timeout 1500 python tools/degree_sweep.py --degrees 4,6,8,3,5,7 --pcg-iters 40 --steps 60 > gpurun_out/r2t_sweep_1gpu.jsonl 2> gpurun_out/r2t_sweep.err
cut -c1-260 gpurun_out/r2t_sweep_1gpu.jsonl; tail -2 gpurun_out/r2t_sweep.err
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2t_pytest_gpu.log
