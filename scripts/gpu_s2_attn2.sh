#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/gpu_tc_attn.py attn2 > gpurun_out/s3_attn2_unit.log 2>&1; echo "exit=$?" >> gpurun_out/s3_attn2_unit.log
cat gpurun_out/s3_attn2_unit.log | tail -20
timeout 300 python tests/test_gpu_chain.py > gpurun_out/s3_chain_unit.log 2>&1; tail -8 gpurun_out/s3_chain_unit.log
timeout 600 python -m pytest tests/test_gpu_tc_arm.py -q -k "loops or attention2 or forward" > gpurun_out/s3_pytest.log 2>&1; tail -8 gpurun_out/s3_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_chain -c 2 -o gpurun_out/s3_chain_prof python tests/test_gpu_chain.py sa_out_q out_ffn1 > gpurun_out/s3_ncu_chain.log 2>&1; tail -3 gpurun_out/s3_ncu_chain.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench_attn2.json 2> gpurun_out/s3_bench_attn2.err; tail -c 900 gpurun_out/s3_bench_attn2.json
echo done
