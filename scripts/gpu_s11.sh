#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python scripts/gpu_attn_trace.py 21 > gpurun_out/s11_attn_trace.log 2>&1; tail -24 gpurun_out/s11_attn_trace.log
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s11_chain_trace.log 2>&1; grep -v trace gpurun_out/s11_chain_trace.log | tail -7
timeout 600 python -m pytest tests/test_gpu_tc_arm.py tests/test_gpu_chain.py -q -k "attention2 or loops or chain" > gpurun_out/s11_pytest.log 2>&1; tail -3 gpurun_out/s11_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err; tail -c 500 gpurun_out/s11_bench.json; tail -3 gpurun_out/s11_bench.err
echo done
