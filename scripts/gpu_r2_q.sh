#!/bin/bash
# round 2, session 2: chain N split + late PDL trigger -- unit tests, per-launch times, whole GPU suite, loop A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -m gpu -q --maxfail=8 > gpurun_out/r2q_pytest_chain.log 2>&1; tail -4 gpurun_out/r2q_pytest_chain.log
timeout 240 python tests/test_gpu_chain.py out_ffn1 ffn2_qkv sa_out_q 2>&1 | grep -v Warn | cut -c1-160 > gpurun_out/r2q_chain_times.txt; tail -40 gpurun_out/r2q_chain_times.txt
timeout 600 python -m pytest tests -m gpu -q --maxfail=8 --deselect tests/test_gpu_chain.py > gpurun_out/r2q_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2q_pytest_gpu.log
A2P_PDL=1 timeout 400 python -m pytest tests/test_gpu_tc_arm.py tests/test_gpu_parity.py -m gpu -q --maxfail=8 -k "units or golden or deterministic" > gpurun_out/r2q_pytest_pdl.log 2>&1; tail -3 gpurun_out/r2q_pytest_pdl.log
short() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["config3_strong"]["value"],1) if d.get("config3_strong") else None, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print(f, "ERR", e)
PY
}
run() { # name, extra bench args, env...
  local name=$1; shift; local args=$1; shift
  env "$@" timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline $args > gpurun_out/r2q_bench_$name.json 2> gpurun_out/r2q_bench_$name.err
  short gpurun_out/r2q_bench_$name.json; tail -1 gpurun_out/r2q_bench_$name.err | cut -c1-160
}
run base "" A2P_CHAIN_NSPLIT=0 A2P_PDL=0
run nsplit "--no-config3" A2P_CHAIN_NSPLIT=1 A2P_PDL=0
run pdl "--no-config3" A2P_CHAIN_NSPLIT=0 A2P_PDL=1
run both "" A2P_CHAIN_NSPLIT=1 A2P_PDL=1
run b4_base "--no-config3 --batch 4" A2P_CHAIN_NSPLIT=0 A2P_PDL=0
run b4_both "--no-config3 --batch 4" A2P_CHAIN_NSPLIT=1 A2P_PDL=1
run b4_nsplit "--no-config3 --batch 4" A2P_CHAIN_NSPLIT=1 A2P_PDL=0
