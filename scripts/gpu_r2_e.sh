#!/bin/bash
# round 2, call E: independent per-group pipelines (A/B against per-step joins), group-count sweep, full test suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2e_pytest.log 2>&1; tail -6 gpurun_out/r2e_pytest.log
short() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["config3_strong"]["value"],1) if d.get("config3_strong") else None, d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["gpu_launches"])
except Exception as e: print(f, "ERR", e)
PY
}
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline"
A2P_GROUP_PIPELINES=0 $B > gpurun_out/r2e_join.json 2> gpurun_out/r2e_join.err; short gpurun_out/r2e_join.json
$B > gpurun_out/r2e_pipes.json 2> gpurun_out/r2e_pipes.err; short gpurun_out/r2e_pipes.json; tail -2 gpurun_out/r2e_pipes.err
A2P_BRANCH_GROUPS=4 $B > gpurun_out/r2e_pipes_g4.json 2> gpurun_out/r2e_pipes_g4.err; short gpurun_out/r2e_pipes_g4.json
# batch 32 with 2 and 4 independent groups (default there is 1 group)
for g in 2 4; do A2P_BRANCH_GROUPS=$g $B --batch 32 --no-config3 --steps 2 > gpurun_out/r2e_b32_g$g.json 2> gpurun_out/r2e_b32_g$g.err; short gpurun_out/r2e_b32_g$g.json; done
for g in 1 2; do A2P_BRANCH_GROUPS=$g $B --batch 16 --no-config3 --steps 2 > gpurun_out/r2e_b16_g$g.json 2> gpurun_out/r2e_b16_g$g.err; short gpurun_out/r2e_b16_g$g.json; done
