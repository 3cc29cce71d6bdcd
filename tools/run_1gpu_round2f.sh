#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ipdg.py -q -m gpu > gpurun_out/r2v_pytest_ipdg.log 2>&1
tail -25 gpurun_out/r2v_pytest_ipdg.log | cut -c1-300
for cfg in "5 44 8" "3 72 8"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ax_hex3d_chain -s 2 -c 1 -f \
    -o gpurun_out/r2v_ncu_chain_n$1 python tools/chain_prof.py --degree $1 --elements $2 --chain $3 --stages 1 --reps 4 \
    > gpurun_out/r2v_ncu_chain_n$1.log 2>&1
  tail -1 gpurun_out/r2v_ncu_chain_n$1.log
done
