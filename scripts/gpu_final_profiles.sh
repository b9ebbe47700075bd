#!/bin/bash
# evidence for the round: GPU suite, default bench line, batch-32 line, ncu launch list of the bench command, full captures of the dominant kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; tail -6 gpurun_out/f_pytest.log
timeout 600 python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err; tail -c 1500 gpurun_out/f_bench_default.json; tail -3 gpurun_out/f_bench_default.err
timeout 600 python bench.py --batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_b32.json 2> gpurun_out/f_bench_b32.err; tail -c 300 gpurun_out/f_bench_b32.json | head -c 300; tail -2 gpurun_out/f_bench_b32.err
# launch list: one 20-step loop of the benchmark workload (same shapes / kernels as the 1000-step loop; every step is identical)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 20 > gpurun_out/f_ncu_launches.log 2>&1; tail -2 gpurun_out/f_ncu_launches.log | cut -c1-200
# full captures from the same process: attention (self / audio cross launches) and chain kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_attn2 -s 96 -c 4 -o gpurun_out/f_attn2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 10 > gpurun_out/f_ncu_attn2.log 2>&1; tail -2 gpurun_out/f_ncu_attn2.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_chain -s 100 -c 8 -o gpurun_out/f_chain python bench.py --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 10 > gpurun_out/f_ncu_chain.log 2>&1; tail -2 gpurun_out/f_ncu_chain.log | cut -c1-200
echo done
