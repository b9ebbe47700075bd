#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/gpu_tc_attn.py lo 2>&1 | grep -v Warn | cut -c1-110 > gpurun_out/r2i_attn_lo.txt; cat gpurun_out/r2i_attn_lo.txt
