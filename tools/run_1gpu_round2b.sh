#!/bin/bash
# round-2 follow-up on one B200: trilinear operator mode (tests + GDOF/s), ogs setup timing, full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_setup_kernels.py -x -q -m gpu > gpurun_out/r2q_pytest_setup_kernels.log 2>&1
tail -3 gpurun_out/r2q_pytest_setup_kernels.log
for lam in 0 1; do
  timeout 600 python tools/chain_tune.py --degree 7 --elements 64 --lam $lam --grid 0:1,4:1 >> gpurun_out/r2q_trilinear_tune.jsonl 2>gpurun_out/r2q_tri_err.log
  timeout 600 python tools/chain_tune.py --degree 7 --elements 64 --lam $lam --trilinear --grid 0:1,4:1,8:1,2:1 >> gpurun_out/r2q_trilinear_tune.jsonl 2>>gpurun_out/r2q_tri_err.log
done
for N in 3 5; do
  timeout 600 python tools/chain_tune.py --degree $N --elements 96 --lam 0 --trilinear --grid 0:1,8:1 >> gpurun_out/r2q_trilinear_tune.jsonl 2>>gpurun_out/r2q_tri_err.log
done
cat gpurun_out/r2q_trilinear_tune.jsonl | cut -c1-400
LIBP_OGS_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench_1gpu.json 2> gpurun_out/r2q_bench_1gpu.err
grep -i "ogs\|setup" gpurun_out/r2q_bench_1gpu.err | head -40
cat gpurun_out/r2q_bench_1gpu.json | cut -c1-1500
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2q_pytest_gpu.log
bash tools/profile_chain_orders.sh
