#!/bin/bash
# round 2, call F: persistent attention CTAs (unit tests + A/B timings + bench), 1000-step goldens (varied and constant guidance)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_arm.py -m gpu -q --maxfail=20 > gpurun_out/r2f_pytest.log 2>&1; tail -5 gpurun_out/r2f_pytest.log
timeout 300 python scripts/gpu_tc_attn.py persist 2>&1 | grep -v Warn | cut -c1-110 > gpurun_out/r2f_attn_persist.txt; cat gpurun_out/r2f_attn_persist.txt
for n in pose_full_b4 pose_full_b4_g2; do timeout 600 python scripts/gpu_loop1000_arms.py $n 2>&1 | grep -v Warn > gpurun_out/r2f_loop1000_$n.txt; tail -8 gpurun_out/r2f_loop1000_$n.txt; done
short() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["config3_strong"]["value"],1) if d.get("config3_strong") else None, d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["roofline"]["forward_ms_by_kernel"])
except Exception as e: print(f, "ERR", e)
PY
}
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline"
for pz in 0 1; do A2P_ATTN_PERSIST=$pz $B > gpurun_out/r2f_persist$pz.json 2> gpurun_out/r2f_persist$pz.err; short gpurun_out/r2f_persist$pz.json; tail -2 gpurun_out/r2f_persist$pz.err; done
A2P_ATTN_PERSIST=1 timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -k "forward_vs or loops or concurrent or rows_independent or benchmarked" > gpurun_out/r2f_pytest_persist.log 2>&1; tail -4 gpurun_out/r2f_pytest_persist.log
