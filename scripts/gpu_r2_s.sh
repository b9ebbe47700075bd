#!/bin/bash
# round 2, session 2: native conditioning encoders (N1), E_B accumulator prefetch, split policy by K0 -- whole GPU suite + loop checks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2s_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2s_pytest_gpu.log
timeout 240 python tests/test_gpu_chain.py out_ffn1 ffn2_qkv sa_out_q 2>&1 | grep -v Warn | grep "nsplit0\|M=2400 mode1" | sed 's/{.*}//' > gpurun_out/r2s_chain_times.txt; head -24 gpurun_out/r2s_chain_times.txt
short() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["config3_strong"]["value"],1) if d.get("config3_strong") else None, d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d.get("one_time",{}).get("conditioning_ms"))
except Exception as e: print(f, "ERR", e)
PY
}
run() { # name, extra bench args, env...
  local name=$1; shift; local args=$1; shift
  env "$@" timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline $args > gpurun_out/r2s_bench_$name.json 2> gpurun_out/r2s_bench_$name.err
  short gpurun_out/r2s_bench_$name.json; tail -1 gpurun_out/r2s_bench_$name.err | cut -c1-160
}
run default "" A2P_DUMMY=1
run b4 "--no-config3 --batch 4" A2P_DUMMY=1
run b32_g2 "--no-config3 --batch 32" A2P_BRANCH_GROUPS=2
run cond_torch "--no-config3" A2P_COND_TORCH=1
