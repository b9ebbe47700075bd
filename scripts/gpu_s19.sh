#!/bin/bash
# per-chunk x-tile stores in the chain kernel + CFG branches as two concurrent forwards (stagger sweep)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
A2P_CHAIN_TRACE=1 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s19_chain_trace.log 2>&1; grep -v trace gpurun_out/s19_chain_trace.log | tail -8; grep trace gpurun_out/s19_chain_trace.log | sed -n '2p' | cut -c1-420
A2P_CHAIN_CLUSTER=2 timeout 300 python tests/test_gpu_chain.py > gpurun_out/s19_chain_cluster.log 2>&1; grep -v trace gpurun_out/s19_chain_cluster.log | tail -8
B="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --diffusion-steps 100"
short() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],2))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
A2P_NO_BRANCH_STREAMS=1 $B > gpurun_out/s19_sweep_off.json 2>gpurun_out/s19_sweep_off.err; short gpurun_out/s19_sweep_off.json
for s in 0 1 2 3 4 6 9; do A2P_BRANCH_STAGGER=$s $B > gpurun_out/s19_sweep_$s.json 2>gpurun_out/s19_sweep_$s.err; short gpurun_out/s19_sweep_$s.json; done
tail -3 gpurun_out/s19_sweep_1.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s19_pytest.log 2>&1; tail -5 gpurun_out/s19_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s19_bench.json 2> gpurun_out/s19_bench.err; tail -c 600 gpurun_out/s19_bench.json; tail -3 gpurun_out/s19_bench.err
echo done
