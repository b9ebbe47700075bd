#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_arm.py -m gpu -q --maxfail=10 -k "attention" > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
timeout 300 python scripts/gpu_tc_attn.py persist 2>&1 | grep -v Warn | cut -c1-110 > gpurun_out/r2k_attn.txt; cat gpurun_out/r2k_attn.txt
for sk in 0 500; do A2P_ATTN_SKEW_NS=$sk timeout 200 python scripts/gpu_attn_trace.py 21 2>&1 | grep -v Warn > gpurun_out/r2k_attn_trace_skew$sk.txt; cat gpurun_out/r2k_attn_trace_skew$sk.txt | cut -c1-230 | tail -16; done
short() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["config3_strong"]["value"],1) if d.get("config3_strong") else None, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print(f, "ERR", e)
PY
}
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; short gpurun_out/r2k_bench.json; tail -2 gpurun_out/r2k_bench.err
