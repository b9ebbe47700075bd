#!/bin/bash
# programmatic dependent launch on the current kernels: A/B on the 100-step loop + GPU suite under A2P_PDL=1
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --diffusion-steps 100"
short() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],2))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
$B > gpurun_out/s22_off.json 2>gpurun_out/s22_off.err; short gpurun_out/s22_off.json
A2P_PDL=1 $B > gpurun_out/s22_pdl.json 2>gpurun_out/s22_pdl.err; short gpurun_out/s22_pdl.json; tail -2 gpurun_out/s22_pdl.err
A2P_PDL=1 A2P_NO_BRANCH_STREAMS=1 $B > gpurun_out/s22_pdl_single.json 2>gpurun_out/s22_pdl_single.err; short gpurun_out/s22_pdl_single.json
A2P_PDL=1 A2P_NO_SIDE_STREAM=1 $B > gpurun_out/s22_pdl_noside.json 2>/dev/null; short gpurun_out/s22_pdl_noside.json
A2P_PDL=1 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s22_pytest_pdl.log 2>&1; tail -5 gpurun_out/s22_pytest_pdl.log
echo done
