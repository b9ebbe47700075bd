#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A2P_CHAIN_CLUSTER=2 A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s16_chain_cl2.log 2>&1; grep -v trace gpurun_out/s16_chain_cl2.log | tail -8; grep trace gpurun_out/s16_chain_cl2.log | sed -n '2p;4p'
A2P_CHAIN_CLUSTER=2 timeout 600 python -m pytest tests/test_gpu_tc_arm.py tests/test_gpu_chain.py tests/test_gpu_parity.py -q -k "loops or chain or golden" > gpurun_out/s16_pytest_cl2.log 2>&1; tail -3 gpurun_out/s16_pytest_cl2.log
A2P_CHAIN_CLUSTER=2 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s16_bench_cl2.json 2> gpurun_out/s16_bench_cl2.err; tail -c 400 gpurun_out/s16_bench_cl2.json; tail -3 gpurun_out/s16_bench_cl2.err
echo done
