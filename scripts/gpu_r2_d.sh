#!/bin/bash
# round 2, call D: attention per-head MMA scheduling A/B (unit timings + correctness + bench), error-source probe of the arms
mkdir -p gpurun_out
for cfg in "0 0" "1 0" "1 300" "1 600" "1 900"; do set -- $cfg
  echo "== decouple=$1 skew=$2" >> gpurun_out/r2d_attn_decouple.txt
  A2P_ATTN_DECOUPLE=$1 A2P_ATTN_SKEW_NS=$2 timeout 300 python scripts/gpu_tc_attn.py decouple 2>&1 | grep -v Warn | cut -c1-110 >> gpurun_out/r2d_attn_decouple.txt
done
cat gpurun_out/r2d_attn_decouple.txt
A2P_ATTN_DECOUPLE=1 timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -k "attention or forward_vs or loops or concurrent or plms or benchmarked" > gpurun_out/r2d_pytest_decouple.log 2>&1; tail -5 gpurun_out/r2d_pytest_decouple.log
timeout 600 python scripts/gpu_arm_error_probe.py > gpurun_out/r2d_arm_error_probe.txt 2>&1; cat gpurun_out/r2d_arm_error_probe.txt | tail -9
for d in 0 1; do
A2P_ATTN_DECOUPLE=$d timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2d_bench_dec$d.json 2> gpurun_out/r2d_bench_dec$d.err
echo "bench decouple=$d rc=$?"; python - gpurun_out/r2d_bench_dec$d.json <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), d["config3_strong"]["value"] if d.get("config3_strong") else None, d["clocks"], d["roofline"]["forward_ms_by_kernel"])
except Exception as e: print(f, "ERR", e)
PY
tail -2 gpurun_out/r2d_bench_dec$d.err
done
