#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s6_chain_trace.log 2>&1; grep -v "trace" gpurun_out/s6_chain_trace.log | tail -8
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s6_pytest.log 2>&1; tail -5 gpurun_out/s6_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; tail -c 600 gpurun_out/s6_bench.json; tail -3 gpurun_out/s6_bench.err
A2P_PDL=1 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s6_bench_pdl.json 2> gpurun_out/s6_bench_pdl.err; tail -c 600 gpurun_out/s6_bench_pdl.json; tail -3 gpurun_out/s6_bench_pdl.err
A2P_PROFILE_DUMP=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 10 > /dev/null 2> gpurun_out/s6_profdump.txt
echo done
