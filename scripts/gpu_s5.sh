#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s5_chain_trace.log 2>&1; tail -14 gpurun_out/s5_chain_trace.log
timeout 300 python scripts/gpu_tc_attn.py attn2poly > gpurun_out/s5_attn2poly.log 2>&1; tail -8 gpurun_out/s5_attn2poly.log
timeout 600 python -m pytest tests/test_gpu_tc_arm.py -q -k "attention2 or loops" > gpurun_out/s5_pytest.log 2>&1; tail -4 gpurun_out/s5_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; tail -c 700 gpurun_out/s5_bench.json; tail -3 gpurun_out/s5_bench.err
echo done
