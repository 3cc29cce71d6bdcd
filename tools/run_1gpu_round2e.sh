#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ipdg.py -x -q -m gpu > gpurun_out/r2u_pytest_ipdg.log 2>&1
tail -25 gpurun_out/r2u_pytest_ipdg.log
timeout 900 python tools/ipdg_bench.py --degree 7 --elements 48 > gpurun_out/r2u_ipdg_bench.jsonl 2> gpurun_out/r2u_ipdg_bench.err
timeout 900 python tools/ipdg_bench.py --degree 3 --elements 96 >> gpurun_out/r2u_ipdg_bench.jsonl 2>> gpurun_out/r2u_ipdg_bench.err
cat gpurun_out/r2u_ipdg_bench.jsonl; tail -3 gpurun_out/r2u_ipdg_bench.err
