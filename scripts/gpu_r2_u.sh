#!/bin/bash
# round 2, session 2: TMA-store E_B epilogue of the chain kernel -- unit tests, timeline, per-launch times, loop A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -m gpu -q --maxfail=8 > gpurun_out/r2u_pytest_chain.log 2>&1; tail -4 gpurun_out/r2u_pytest_chain.log
bash scripts/gpu_r2_t.sh > /dev/null 2>&1; cp gpurun_out/r2t_chain_trace.txt gpurun_out/r2u_chain_trace_tma.txt; cut -c1-400 gpurun_out/r2u_chain_trace_tma.txt
timeout 240 python tests/test_gpu_chain.py out_ffn1 ffn2_qkv sa_out_q 2>&1 | grep -v Warn | grep "nsplit0\|M=2400 mode1" | sed 's/{.*}//' > gpurun_out/r2u_chain_times.txt; head -12 gpurun_out/r2u_chain_times.txt
timeout 600 python -m pytest tests -m gpu -q --maxfail=8 --deselect tests/test_gpu_chain.py > gpurun_out/r2u_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2u_pytest_gpu.log
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
  env "$@" timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline $args > gpurun_out/r2u_bench_$name.json 2> gpurun_out/r2u_bench_$name.err
  short gpurun_out/r2u_bench_$name.json; tail -1 gpurun_out/r2u_bench_$name.err | cut -c1-160
}
run tma "" A2P_CHAIN_EB_TMA=1
run old "" A2P_CHAIN_EB_TMA=0
