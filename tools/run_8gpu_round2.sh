TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/r2j_pytest_8gpu.log 2>&1; tail -12 gpurun_out/r2j_pytest_8gpu.log | cut -c1-250
$TR bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r2j_bench_8gpu.json 2> gpurun_out/r2j_8gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_8gpu.json')); print('C2/C3 8gpu', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac'], d['pcg']['value'], d['pcg']['ms_per_iteration'], d['pcg']['solve_e2e'])"
$TR bench.py --gpus 8 --workload c4 > gpurun_out/r2j_bench_c4_8gpu.json 2>> gpurun_out/r2j_8gpu.err; cut -c1-700 gpurun_out/r2j_bench_c4_8gpu.json
$TR tools/degree_sweep.py --degrees 3,4,5,6,7,8 --pcg-iters 40 --steps 60 > gpurun_out/r2j_sweep_8gpu.jsonl 2>> gpurun_out/r2j_8gpu.err; cut -c1-300 gpurun_out/r2j_sweep_8gpu.jsonl
tail -5 gpurun_out/r2j_8gpu.err
