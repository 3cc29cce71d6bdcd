#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py tests/test_gpu_setup_kernels.py tests/test_gpu_multigrid.py -x -q -m gpu > gpurun_out/r2w_pytest_chain_parity.log 2>&1
tail -5 gpurun_out/r2w_pytest_chain_parity.log
timeout 1500 python tools/degree_sweep.py --degrees 3,5,4,6,8,7 --pcg-iters 40 --steps 60 > gpurun_out/r2w_sweep_1gpu.jsonl 2> gpurun_out/r2w_sweep.err
cut -c1-260 gpurun_out/r2w_sweep_1gpu.jsonl; tail -2 gpurun_out/r2w_sweep.err
for cfg in "5 44 8" "3 72 8"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ax_hex3d_chain -s 2 -c 1 -f \
    -o gpurun_out/r2w_ncu_chain_dealt_n$1 python tools/chain_prof.py --degree $1 --elements $2 --chain $3 --stages 1 --reps 4 \
    > gpurun_out/r2w_ncu_chain_dealt_n$1.log 2>&1
  tail -1 gpurun_out/r2w_ncu_chain_dealt_n$1.log
done
