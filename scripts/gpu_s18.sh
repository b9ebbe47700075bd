#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s18_chain_trace.log 2>&1; grep -v trace gpurun_out/s18_chain_trace.log | tail -8; grep trace gpurun_out/s18_chain_trace.log | sed -n '2p;8p' | cut -c1-420
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s18_pytest.log 2>&1; tail -5 gpurun_out/s18_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s18_bench.json 2> gpurun_out/s18_bench.err; tail -c 500 gpurun_out/s18_bench.json; tail -3 gpurun_out/s18_bench.err
echo done
