python -m pytest tests -x -q -m gpu > gpurun_out/r2p_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2p_pytest_gpu.log
LIBP_OGS_TIMING=1 python bench.py --mode 0 --no-cpu --no-e2e --steps 50 > gpurun_out/r2p_bench_mode0.json 2> gpurun_out/r2p_bench_mode0.err; grep "ogs setup" gpurun_out/r2p_bench_mode0.err | head -12; python -c "
import json
for l in open('gpurun_out/r2p_bench_mode0.json'):
    if l.startswith('{'):
        d=json.loads(l); print('mode0', d['value'], d['ms_per_step'], d['setup_seconds'], d['pcg']['value'] if d.get('pcg') else None)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ax_hex3d_chain|zero_fill" -s 2 -c 2 -f -o gpurun_out/r2p_step_n7_e64 python tools/chain_prof.py --elements 64 --chain 4 --stages 1 --reps 3 > gpurun_out/r2p_ncu.log 2>&1; tail -2 gpurun_out/r2p_ncu.log
python tools/degree_sweep.py --degrees 3,4,5,6,7,8 --pcg-iters 40 --steps 60 --cpu-seconds 6 > gpurun_out/r2p_sweep_1gpu.jsonl 2> gpurun_out/r2p_sweep.err; cut -c1-250 gpurun_out/r2p_sweep_1gpu.jsonl; tail -2 gpurun_out/r2p_sweep.err
