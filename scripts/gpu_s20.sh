#!/bin/bash
# CFG branches as concurrent forwards: chain-kernel launch priority x stagger sweep, split-KV on/off
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
for s in 0 1 2 3 5; do A2P_CHAIN_PRIO=-3 A2P_BRANCH_STAGGER=$s $B > gpurun_out/s20_prio_$s.json 2>gpurun_out/s20_prio_$s.err; short gpurun_out/s20_prio_$s.json; done
tail -2 gpurun_out/s20_prio_0.err
A2P_NO_SPLIT_KV=1 $B > gpurun_out/s20_nosplit.json 2>gpurun_out/s20_nosplit.err; short gpurun_out/s20_nosplit.json
A2P_NO_SPLIT_KV=1 A2P_CHAIN_PRIO=-3 A2P_BRANCH_STAGGER=2 $B > gpurun_out/s20_nosplit_p2.json 2>/dev/null; short gpurun_out/s20_nosplit_p2.json
$B > gpurun_out/s20_default.json 2>/dev/null; short gpurun_out/s20_default.json
echo done
