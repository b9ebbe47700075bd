#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s15_chain_trace.log 2>&1; grep -v trace gpurun_out/s15_chain_trace.log | tail -7; grep trace gpurun_out/s15_chain_trace.log | sed -n '2p;8p'
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s15_pytest.log 2>&1; tail -4 gpurun_out/s15_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s15_bench.json 2> gpurun_out/s15_bench.err; tail -c 400 gpurun_out/s15_bench.json; tail -3 gpurun_out/s15_bench.err
timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 32 > gpurun_out/s15_bench_b32.json 2> gpurun_out/s15_bench_b32.err; tail -c 700 gpurun_out/s15_bench_b32.json; tail -3 gpurun_out/s15_bench_b32.err
echo done
