#!/bin/bash
# short-key-set attention kernel (keyframe cross-attention): unit tests, A/B on the 100-step loop, full GPU suite, bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_arm.py -m gpu -q -x -k "attention_short" > gpurun_out/s21_short_unit.log 2>&1; tail -5 gpurun_out/s21_short_unit.log
B="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --diffusion-steps 100"
short() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],2), d["roofline"]["forward_ms_by_kernel"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
A2P_NO_ATTN_SHORT=1 $B > gpurun_out/s21_off.json 2>gpurun_out/s21_off.err; short gpurun_out/s21_off.json
$B > gpurun_out/s21_on.json 2>gpurun_out/s21_on.err; short gpurun_out/s21_on.json; tail -2 gpurun_out/s21_on.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s21_pytest.log 2>&1; tail -5 gpurun_out/s21_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err; short gpurun_out/s21_bench.json; tail -3 gpurun_out/s21_bench.err
echo done
