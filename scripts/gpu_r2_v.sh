#!/bin/bash
# round 2, session 2: face / pipeline workloads on the final code (PDL default), N split A/B with the TMA-store epilogue
mkdir -p gpurun_out
short() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1) if d.get("e2e") else None, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print(f, "ERR", e)
PY
}
run() { # name, extra bench args, env...
  local name=$1; shift; local args=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline $args > gpurun_out/r2v_bench_$name.json 2> gpurun_out/r2v_bench_$name.err
  short gpurun_out/r2v_bench_$name.json; tail -1 gpurun_out/r2v_bench_$name.err | cut -c1-160
}
run nsplit0 "--no-config3" A2P_CHAIN_NSPLIT=0
run nsplit1 "--no-config3" A2P_CHAIN_NSPLIT=1
run b4_nsplit0 "--no-config3 --batch 4" A2P_CHAIN_NSPLIT=0
run b4_nsplit1 "--no-config3 --batch 4" A2P_CHAIN_NSPLIT=1
run face "--workload face" A2P_DUMMY=1
run face_nopdl "--workload face" A2P_PDL=0
run pipeline "--workload pipeline" A2P_DUMMY=1
