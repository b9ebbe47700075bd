#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s14_chain_trace.log 2>&1; grep -v trace gpurun_out/s14_chain_trace.log | tail -7; grep trace gpurun_out/s14_chain_trace.log | sed -n '2p;8p'
timeout 600 python -m pytest tests/test_gpu_tc_arm.py tests/test_gpu_chain.py tests/test_gpu_parity.py -q -k "loops or chain or golden" > gpurun_out/s14_pytest.log 2>&1; tail -3 gpurun_out/s14_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err; tail -c 900 gpurun_out/s14_bench.json; tail -3 gpurun_out/s14_bench.err
echo done
