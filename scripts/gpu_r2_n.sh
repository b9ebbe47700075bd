#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_face_arm_probe.py 2>&1 | grep -v Warn > gpurun_out/r2n_face_arm_probe.txt; cat gpurun_out/r2n_face_arm_probe.txt
for t in 3 2; do timeout 900 python bench.py --workload face --split-terms $t --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2n_bench_face_terms$t.json 2> gpurun_out/r2n_bench_face_terms$t.err; python - gpurun_out/r2n_bench_face_terms$t.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), d["ms_per_step"], d["roofline"]["forward_ms_by_kernel"])
except Exception as e: print("ERR", e)
PY
done
