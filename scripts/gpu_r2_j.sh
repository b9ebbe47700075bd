#!/bin/bash
mkdir -p gpurun_out
for sk in 0 300 700; do A2P_ATTN_SKEW_NS=$sk timeout 200 python scripts/gpu_attn_trace.py 21 2>&1 | grep -v Warn > gpurun_out/r2j_attn_trace_skew$sk.txt; cat gpurun_out/r2j_attn_trace_skew$sk.txt | cut -c1-230; done
