#!/bin/bash
# CFG step cut into 2 G concurrent forwards (branches x groups of batch rows): equivalence test, timing sweep, GPU suite
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_arm.py -m gpu -q -x -k "concurrent_forward" > gpurun_out/s23_units_test.log 2>&1; tail -6 gpurun_out/s23_units_test.log
B="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --diffusion-steps 100"
short() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],2), d["gpu_launches"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
for g in 1 2 4; do A2P_BRANCH_GROUPS=$g $B > gpurun_out/s23_g$g.json 2>gpurun_out/s23_g$g.err; short gpurun_out/s23_g$g.json; done
tail -2 gpurun_out/s23_g4.err
for g in 2 4; do A2P_BRANCH_GROUPS=$g $B --batch 32 > gpurun_out/s23_b32_g$g.json 2>gpurun_out/s23_b32_g$g.err; short gpurun_out/s23_b32_g$g.json; done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s23_pytest.log 2>&1; tail -5 gpurun_out/s23_pytest.log
echo done
