#!/bin/bash
# round 2, call B: full GPU suite incl. the 1000-step golden and the CTA-pair chain unit tests; chain unit timings in both
# modes; default bench line; bench with the pair-mode chain kernels; per-launch profile at B=32
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -12 gpurun_out/r2b_pytest.log
timeout 600 python tests/test_gpu_chain.py > gpurun_out/r2b_chain_unit.txt 2>&1; cat gpurun_out/r2b_chain_unit.txt | tail -14
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
A2P_CHAIN_PAIR=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2b_bench_pair.json 2> gpurun_out/r2b_bench_pair.err
echo "bench pair rc=$?"; tail -c 1500 gpurun_out/r2b_bench_pair.json; tail -5 gpurun_out/r2b_bench_pair.err
A2P_PROFILE_DUMP=1 timeout 600 python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-config3 > gpurun_out/r2b_b32.json 2> gpurun_out/r2b_b32_prof.txt
echo "b32 rc=$?"
A2P_CHAIN_PAIR=1 A2P_PROFILE_DUMP=1 timeout 600 python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-config3 > gpurun_out/r2b_b32_pair.json 2> gpurun_out/r2b_b32_pair_prof.txt
echo "b32 pair rc=$?"
