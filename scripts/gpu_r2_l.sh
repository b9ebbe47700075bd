#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_arm.py -m gpu -q --maxfail=10 -k "attention" > gpurun_out/r2l_pytest.log 2>&1; tail -3 gpurun_out/r2l_pytest.log
timeout 300 python scripts/gpu_tc_attn.py persist 2>&1 | grep -v Warn | cut -c1-110 > gpurun_out/r2l_attn.txt; tail -7 gpurun_out/r2l_attn.txt
short() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["config3_strong"]["value"],1) if d.get("config3_strong") else None, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print(f, "ERR", e)
PY
}
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; short gpurun_out/r2l_bench.json; tail -2 gpurun_out/r2l_bench.err
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -k "forward_vs or loops or concurrent or rows_independent or benchmarked or caller or variants" > gpurun_out/r2l_pytest2.log 2>&1; tail -3 gpurun_out/r2l_pytest2.log
