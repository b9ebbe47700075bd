#!/bin/bash
# round 2, call G: defaults (persistent attention, 4-chain max) vs + chain pair auto mode; attention unit timings
mkdir -p gpurun_out
timeout 300 python scripts/gpu_tc_attn.py persist 2>&1 | grep -v Warn | cut -c1-110 > gpurun_out/r2g_attn.txt; cat gpurun_out/r2g_attn.txt
short() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["config3_strong"]["value"],1) if d.get("config3_strong") else None, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print(f, "ERR", e)
PY
}
B="timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline"
for pair in 0 2 0 2; do A2P_CHAIN_PAIR=$pair $B > gpurun_out/r2g_pair$pair.json 2> gpurun_out/r2g_pair$pair.err; short gpurun_out/r2g_pair$pair.json; tail -1 gpurun_out/r2g_pair$pair.err; done
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -k "attention2 or forward_vs or loops or benchmarked" > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
