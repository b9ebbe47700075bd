#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gpu_mma_rate.py tmem 2>&1 | grep -v Warn > gpurun_out/r2m_tmem_ldst_rate.txt; cat gpurun_out/r2m_tmem_ldst_rate.txt
