#!/bin/bash
# round 2, call C: CTA-pair chain kernel after the m_full fix (unit tests, bench at B=8 incl. config3), the 1000-step golden on
# every arithmetic arm, face workload line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -m gpu -q --maxfail=20 > gpurun_out/r2c_pytest.log 2>&1; tail -4 gpurun_out/r2c_pytest.log
timeout 900 python scripts/gpu_loop1000_arms.py > gpurun_out/r2_loop1000_arms.txt 2>&1; grep -v Warn gpurun_out/r2_loop1000_arms.txt | tail -6
A2P_CHAIN_PAIR=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c_bench_pair.json 2> gpurun_out/r2c_bench_pair.err
echo "bench pair rc=$?"; python - <<'PY'
import json
for f in ["gpurun_out/r2c_bench_pair.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["e2e"]["value"],1), d["config3_strong"]["value"] if d.get("config3_strong") else None, d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r2c_bench_pair.err
A2P_CHAIN_PAIR=1 timeout 600 python -m pytest tests/test_gpu_tc_arm.py -m gpu -q -k "loops or concurrent or forward_vs" > gpurun_out/r2c_pytest_pair.log 2>&1; tail -3 gpurun_out/r2c_pytest_pair.log
timeout 900 python bench.py --workload face --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_bench_face.json 2> gpurun_out/r2c_bench_face.err
echo "face rc=$?"; tail -c 1200 gpurun_out/r2c_bench_face.json; tail -3 gpurun_out/r2c_bench_face.err
