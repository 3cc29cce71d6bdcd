#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR tools/trace_pcg.py --elements 64 --iters 6 --tag r2y_2gpu_e64 > gpurun_out/r2y_trace.log 2>&1; tail -2 gpurun_out/r2y_trace.log
python tools/trace_summary.py gpurun_out/r2y_2gpu_e64_trace.json > gpurun_out/r2y_timeline_2gpu_e64.txt 2>&1; head -40 gpurun_out/r2y_timeline_2gpu_e64.txt
timeout 300 $TR tools/trace_pcg.py --elements 32 --iters 6 --tag r2y_2gpu_e32 >> gpurun_out/r2y_trace.log 2>&1
python tools/trace_summary.py gpurun_out/r2y_2gpu_e32_trace.json > gpurun_out/r2y_timeline_2gpu_e32.txt 2>&1; head -30 gpurun_out/r2y_timeline_2gpu_e32.txt
rm -f gpurun_out/*_chrome.json
timeout 300 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/r2y_pytest_2gpu.log 2>&1; tail -3 gpurun_out/r2y_pytest_2gpu.log | cut -c1-200
