#!/bin/bash
# round 2, session 2: defaults sweep with N split + late PDL (row groups, min halves per part, CTA budget), B = 16 point
mkdir -p gpurun_out
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
  env "$@" timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-config3 $args > gpurun_out/r2r_bench_$name.json 2> gpurun_out/r2r_bench_$name.err
  short gpurun_out/r2r_bench_$name.json; tail -1 gpurun_out/r2r_bench_$name.err | cut -c1-160
}
run both "" A2P_CHAIN_NSPLIT=1 A2P_PDL=1
run both_g1 "" A2P_CHAIN_NSPLIT=1 A2P_PDL=1 A2P_BRANCH_GROUPS=1
run both_g4 "" A2P_CHAIN_NSPLIT=1 A2P_PDL=1 A2P_BRANCH_GROUPS=4
run both_minh1 "" A2P_CHAIN_NSPLIT=1 A2P_PDL=1 A2P_CHAIN_SPLIT_MINH=1
run both_minh4 "" A2P_CHAIN_NSPLIT=1 A2P_PDL=1 A2P_CHAIN_SPLIT_MINH=4
run both_budget320 "" A2P_CHAIN_NSPLIT=1 A2P_PDL=1 A2P_CHAIN_SPLIT_BUDGET=320
run pdl_only "" A2P_CHAIN_NSPLIT=0 A2P_PDL=1
run b16_base "--batch 16" A2P_CHAIN_NSPLIT=0 A2P_PDL=0
run b16_both "--batch 16" A2P_CHAIN_NSPLIT=1 A2P_PDL=1
run b16_both_g1 "--batch 16" A2P_CHAIN_NSPLIT=1 A2P_PDL=1 A2P_BRANCH_GROUPS=1
