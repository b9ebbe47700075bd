#!/bin/bash
# evidence for the round: ncu launch list of the bench command + full captures of the two dominant kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
# launch list: one 20-step loop of the benchmark workload (same shapes / kernels as the 1000-step loop; every step is identical)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 20 > gpurun_out/f_ncu_launches.log 2>&1; tail -2 gpurun_out/f_ncu_launches.log | cut -c1-200
# full captures from the same process: attention (audio cross launch = 2nd attention launch of a layer) and chain kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_attn2 -s 40 -c 3 -o gpurun_out/f_attn2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 10 > gpurun_out/f_ncu_attn2.log 2>&1; tail -2 gpurun_out/f_ncu_attn2.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_chain -s 50 -c 4 -o gpurun_out/f_chain python bench.py --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 10 > gpurun_out/f_ncu_chain.log 2>&1; tail -2 gpurun_out/f_ncu_chain.log | cut -c1-200
echo done
