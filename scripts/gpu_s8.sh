#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/gpu_tc_attn.py attn3 > gpurun_out/s8_attn3.log 2>&1; tail -19 gpurun_out/s8_attn3.log
A2P_ATTN2=5 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s8_bench_attn3.json 2> gpurun_out/s8_bench_attn3.err; tail -c 500 gpurun_out/s8_bench_attn3.json; tail -3 gpurun_out/s8_bench_attn3.err
A2P_ATTN2=5 timeout 600 python -m pytest tests/test_gpu_tc_arm.py tests/test_gpu_parity.py -q -k "loops or golden" > gpurun_out/s8_pytest_attn3.log 2>&1; tail -4 gpurun_out/s8_pytest_attn3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_attn3 -c 1 -o gpurun_out/s8_attn3_prof python scripts/gpu_tc_attn.py prof 24 > gpurun_out/s8_ncu_attn3.log 2>&1; tail -2 gpurun_out/s8_ncu_attn3.log
echo done
