#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s7_chain_trace.log 2>&1; grep -v "trace" gpurun_out/s7_chain_trace.log | tail -8
timeout 300 python scripts/gpu_tc_attn.py attn3 > gpurun_out/s7_attn3.log 2>&1; tail -20 gpurun_out/s7_attn3.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err; tail -c 500 gpurun_out/s7_bench.json; tail -3 gpurun_out/s7_bench.err
A2P_ATTN2=5 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s7_bench_attn3.json 2> gpurun_out/s7_bench_attn3.err; tail -c 500 gpurun_out/s7_bench_attn3.json; tail -3 gpurun_out/s7_bench_attn3.err
A2P_ATTN2=5 timeout 600 python -m pytest tests/test_gpu_tc_arm.py tests/test_gpu_parity.py -q -k "loops or golden" > gpurun_out/s7_pytest_attn3.log 2>&1; tail -4 gpurun_out/s7_pytest_attn3.log
echo done
