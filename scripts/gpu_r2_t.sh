#!/bin/bash
# chain kernel timeline (clock64 of CTA 0) for the FFN1 and FFN2->QKV launches at 19 tiles: who waits for whom in the GEMM1 / E_B phase
mkdir -p gpurun_out
for c in out_ffn1 ffn2_qkv; do
A2P_CHAIN_TRACE=1 timeout 120 python - $c <<'PY' 2>&1 | grep -v Warn | grep "trace\|mode" | cut -c1-1500
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import test_gpu_chain as t
for mode in (1,):
    res, _ = t.run_case(sys.argv[1], mode=mode, M=2400)
    print("mode", mode, sys.argv[1], {k: f"{e:.2e}" for k, (e, s) in res.items()})
PY
done > gpurun_out/r2t_chain_trace.txt 2>&1
cat gpurun_out/r2t_chain_trace.txt
