#!/bin/bash
# compute-sanitizer memcheck over the chain kernel (N split, CTA pairs, TMA-store epilogue, ragged last tile) and one small fused loop
mkdir -p gpurun_out
cat > /tmp/san_chain.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import test_gpu_chain as t
for mode in (1, 2):
    for nsplit in (0, 3):
        for name in ("ragged", "sa_out_q"):
            res, _ = t.run_case(name, mode=mode, nsplit=nsplit, M=None if name == "ragged" else 1200)
            print("mode", mode, "nsplit", nsplit, name, {k: f"{e:.2e}" for k, (e, s) in res.items()}, flush=True)
PY
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_chain.py > gpurun_out/r2w_memcheck_chain.log 2>&1; tail -12 gpurun_out/r2w_memcheck_chain.log | cut -c1-200
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/r2w_memcheck_smoke.log 2>&1; tail -6 gpurun_out/r2w_memcheck_smoke.log | cut -c1-200
