#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tests/test_gpu_chain.py > gpurun_out/s4_chain_unit.log 2>&1; tail -8 gpurun_out/s4_chain_unit.log
timeout 300 python scripts/gpu_tc_attn.py attn2poly > gpurun_out/s4_attn2poly.log 2>&1; tail -8 gpurun_out/s4_attn2poly.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s4_pytest.log 2>&1; tail -8 gpurun_out/s4_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; tail -c 700 gpurun_out/s4_bench.json; tail -3 gpurun_out/s4_bench.err
A2P_PROFILE_DUMP=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --diffusion-steps 10 > /dev/null 2> gpurun_out/s4_profdump.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_attn2 -c 1 -o gpurun_out/s4_attn2_prof python scripts/gpu_tc_attn.py prof 21 > gpurun_out/s4_ncu_attn2.log 2>&1; tail -2 gpurun_out/s4_ncu_attn2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_chain -c 1 -o gpurun_out/s4_chain_prof python tests/test_gpu_chain.py sa_out_q > gpurun_out/s4_ncu_chain.log 2>&1; tail -2 gpurun_out/s4_ncu_chain.log
echo done
