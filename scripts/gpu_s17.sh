#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/gpu_tc_attn.py attn2 > gpurun_out/s17_attn2.log 2>&1; grep "terms=21\|terms=2 " gpurun_out/s17_attn2.log | cut -c1-110
A2P_NO_SPLIT_KV=1 timeout 200 python scripts/gpu_tc_attn.py prof 21 2>&1 | tail -2 | cut -c1-110
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s17_pytest.log 2>&1; tail -5 gpurun_out/s17_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err; tail -c 500 gpurun_out/s17_bench.json; tail -3 gpurun_out/s17_bench.err
echo done
