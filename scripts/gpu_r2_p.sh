#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_arm.py -m gpu -q --maxfail=10 -k "attention2" > gpurun_out/r2p_pytest.log 2>&1; tail -4 gpurun_out/r2p_pytest.log
timeout 300 python scripts/gpu_tc_attn.py sk 2>&1 | grep -v Warn | cut -c1-110 > gpurun_out/r2p_attn_sk.txt; cat gpurun_out/r2p_attn_sk.txt
short() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["config3_strong"]["value"],1) if d.get("config3_strong") else None, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print(f, "ERR", e)
PY
}
for av in 2 8; do A2P_ATTN2=$av timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2p_bench_av$av.json 2> gpurun_out/r2p_bench_av$av.err; short gpurun_out/r2p_bench_av$av.json; tail -1 gpurun_out/r2p_bench_av$av.err | cut -c1-200; done
