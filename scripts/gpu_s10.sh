#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py sa_out_q ffn2_qkv > gpurun_out/s10_chain_trace.log 2>&1; tail -4 gpurun_out/s10_chain_trace.log | cut -c1-400
timeout 300 python scripts/gpu_tc_attn.py attn2poly > gpurun_out/s10_attn2.log 2>&1; tail -7 gpurun_out/s10_attn2.log | cut -c1-120
timeout 600 python -m pytest tests/test_gpu_tc_arm.py tests/test_gpu_chain.py -q -k "attention2 or loops or chain" > gpurun_out/s10_pytest.log 2>&1; tail -3 gpurun_out/s10_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err; tail -c 500 gpurun_out/s10_bench.json; tail -3 gpurun_out/s10_bench.err
echo done
