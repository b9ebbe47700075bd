#!/bin/bash
# session-2 baseline: gpu tests, bench at P=2 / P=3, per-launch dump at P=2, ncu launch list of the bench command
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s2_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/s2_pytest.log
tail -3 gpurun_out/s2_pytest.log
timeout 400 python bench.py --split-terms 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench_p2.json 2> gpurun_out/s2_bench_p2.err; tail -c 600 gpurun_out/s2_bench_p2.json
A2P_PROFILE_DUMP=1 timeout 300 python bench.py --split-terms 2 --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 10 > /dev/null 2> gpurun_out/s2_profdump_p2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 500 --csv --log-file gpurun_out/s2_launches_p2.csv python bench.py --split-terms 2 --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 20 > gpurun_out/s2_ncu_bench.log 2>&1
echo done
