#!/bin/bash
# round 2 evidence: full GPU suite, smoke, default bench line (driver flags), ncu launch list of the bench command, ncu --set full captures of the
# dominant kernels from the bench process (B = 8 shapes and B = 32 shapes)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_default_line.json 2> gpurun_out/r02_bench_default_line.err; tail -c 600 gpurun_out/r02_bench_default_line.json; tail -2 gpurun_out/r02_bench_default_line.err
# launch list: a 20-step loop of the benchmark workload (every diffusion step launches the same kernels)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-config3 --diffusion-steps 20 > gpurun_out/r02_ncu_launches.log 2>&1; tail -2 gpurun_out/r02_ncu_launches.log | cut -c1-200
# full captures, B = 8 launch shapes (4-row forwards): attention + chain
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_attn2 -s 96 -c 4 -o gpurun_out/r02_attn2_b8 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-config3 --diffusion-steps 10 > gpurun_out/r02_ncu_attn2_b8.log 2>&1; tail -1 gpurun_out/r02_ncu_attn2_b8.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_chain -s 100 -c 8 -o gpurun_out/r02_chain_b8 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-config3 --diffusion-steps 10 > gpurun_out/r02_ncu_chain_b8.log 2>&1; tail -1 gpurun_out/r02_ncu_chain_b8.log | cut -c1-200
# full captures, B = 32 launch shapes (config 3: 32-row forwards, persistent attention CTAs)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_attn2 -s 48 -c 4 -o gpurun_out/r02_attn2_b32 python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-config3 --diffusion-steps 10 > gpurun_out/r02_ncu_attn2_b32.log 2>&1; tail -1 gpurun_out/r02_ncu_attn2_b32.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_chain -s 50 -c 8 -o gpurun_out/r02_chain_b32 python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-config3 --diffusion-steps 10 > gpurun_out/r02_ncu_chain_b32.log 2>&1; tail -1 gpurun_out/r02_ncu_chain_b32.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
