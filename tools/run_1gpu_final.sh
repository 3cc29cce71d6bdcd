#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2z_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 600 python bench.py > gpurun_out/r2z_bench_e64.json 2> gpurun_out/r2z_bench.err; cut -c1-300 gpurun_out/r2z_bench_e64.json; tail -2 gpurun_out/r2z_bench.err
