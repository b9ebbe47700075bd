#!/bin/bash
# round 2, call A: full GPU test suite (new parity cases), the default bench line with config3 / gpu / cpu baselines,
# per-launch times at B=32 for kernel planning
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -k "not benchmarked_configuration" > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
A2P_PROFILE_DUMP=1 timeout 600 python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-config3 > gpurun_out/r2a_b32.json 2> gpurun_out/r2a_b32_prof.txt
echo "b32 rc=$?"
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
