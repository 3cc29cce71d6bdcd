#!/bin/bash
# final 8-GPU pass of round 2: multi-GPU check, weak-scaling sweep with the dense / dealt layouts, strong-scaling bench
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 150 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/r2x_pytest_8gpu.log 2>&1; tail -3 gpurun_out/r2x_pytest_8gpu.log | cut -c1-250
timeout 200 $TR tools/degree_sweep.py --degrees 4,5,6,8,3,7 --pcg-iters 40 --steps 60 > gpurun_out/r2x_sweep_8gpu.jsonl 2> gpurun_out/r2x_8gpu.err; cut -c1-300 gpurun_out/r2x_sweep_8gpu.jsonl
timeout 150 $TR bench.py --gpus 8 --steps 200 --warmup 10 --no-e2e > gpurun_out/r2x_bench_8gpu.json 2>> gpurun_out/r2x_8gpu.err; cut -c1-400 gpurun_out/r2x_bench_8gpu.json
tail -3 gpurun_out/r2x_8gpu.err
