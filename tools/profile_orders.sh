#!/bin/bash
# One `ncu --set full` capture of the fused even-odd Ax kernel per order (run on a B200 through gpurun), e.g.
#   gpurun --timeout 600 -- 'bash tools/profile_orders.sh 3 4 5 6 7 8'
# Writes gpurun_out/ax_n<N>.ncu-rep; summarise here with  python tools/ncu_summary.py gpurun_out/ax_n<N>.ncu-rep
# Box edges keep ~4-7 M nodes per launch so the replay passes stay short.
set -u
mkdir -p gpurun_out
declare -A EDGE=([1]=128 [2]=80 [3]=56 [4]=44 [5]=36 [6]=30 [7]=26 [8]=24)
for N in "$@"; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:ax_hex3d_t_kernel -c 1 -f \
    -o gpurun_out/ax_n$N python tools/ax_run.py --degree "$N" --elements "${EDGE[$N]}" --reps 2 > gpurun_out/ax_n$N.log 2>&1
  tail -1 gpurun_out/ax_n$N.log
done
ls -la gpurun_out/ax_n*.ncu-rep
