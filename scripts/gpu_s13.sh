#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s13_chain_trace.log 2>&1; grep -v trace gpurun_out/s13_chain_trace.log | tail -7; grep trace gpurun_out/s13_chain_trace.log | sed -n '2p'
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s13_pytest.log 2>&1; tail -5 gpurun_out/s13_pytest.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err; tail -c 600 gpurun_out/s13_bench.json; tail -3 gpurun_out/s13_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s13_ref.json 2> gpurun_out/s13_ref.err; tail -c 300 gpurun_out/s13_ref.json
python __graft_entry__.py --smoke > gpurun_out/s13_smoke.log 2>&1; tail -2 gpurun_out/s13_smoke.log
echo done
