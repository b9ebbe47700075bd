#!/bin/bash
# chain-kernel bring-up: unit cases (errors + kernel times), gpu tests, bench, per-launch dump
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tests/test_gpu_chain.py > gpurun_out/s2_chain_unit.log 2>&1; echo "unit exit=$?" >> gpurun_out/s2_chain_unit.log
cat gpurun_out/s2_chain_unit.log | tail -12
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s2_pytest2.log 2>&1; echo "pytest exit=$?" >> gpurun_out/s2_pytest2.log
tail -15 gpurun_out/s2_pytest2.log
timeout 400 python bench.py --steps 2 --warmup 3 > gpurun_out/s2_bench_chain.json 2> gpurun_out/s2_bench_chain.err; tail -c 1500 gpurun_out/s2_bench_chain.json; tail -3 gpurun_out/s2_bench_chain.err
A2P_PROFILE_DUMP=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 10 > /dev/null 2> gpurun_out/s2_profdump_chain.txt
echo done
