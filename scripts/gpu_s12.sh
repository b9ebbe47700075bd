#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for sk in 0 300 500 700 1000; do echo "skew=$sk"; A2P_ATTN_SKEW_NS=$sk timeout 200 python scripts/gpu_tc_attn.py prof 21 2>&1 | tail -2 | cut -c1-110; A2P_ATTN_SKEW_NS=$sk timeout 200 python - <<'PY' 2>&1 | tail -3 | cut -c1-110
import sys; sys.path.insert(0, ".")
from scripts.gpu_tc_attn import run
run(21, 16, 600, 256, 32, 1998, 2)
run(21, 16, 600, 256, 32, 600, 0)
PY
done > gpurun_out/s12_attn_skew.log 2>&1
cat gpurun_out/s12_attn_skew.log | grep -v DONE
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s12_chain_trace.log 2>&1; grep -v trace gpurun_out/s12_chain_trace.log | tail -7
timeout 600 python -m pytest tests/test_gpu_tc_arm.py tests/test_gpu_chain.py -q -k "attention2 or loops or chain" > gpurun_out/s12_pytest.log 2>&1; tail -3 gpurun_out/s12_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err; tail -c 400 gpurun_out/s12_bench.json; tail -3 gpurun_out/s12_bench.err
A2P_ATTN_SKEW_NS=500 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s12_bench_skew500.json 2> gpurun_out/s12_bench_skew500.err; tail -c 400 gpurun_out/s12_bench_skew500.json
echo done
