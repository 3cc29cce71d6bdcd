#!/bin/bash
# ncu --set full of the element-chain kernel at the orders that sit below 75 % of the roofline (one capture per order)
mkdir -p gpurun_out
for cfg in "4 56 8" "6 38 8" "8 28 4"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ax_hex3d_chain -s 2 -c 1 -f \
    -o gpurun_out/r2r_ncu_chain_n$1 python tools/chain_prof.py --degree $1 --elements $2 --chain $3 --stages 1 --reps 4 \
    > gpurun_out/r2r_ncu_chain_n$1.log 2>&1
  tail -2 gpurun_out/r2r_ncu_chain_n$1.log
done
ls -la gpurun_out/*.ncu-rep
