#!/bin/bash
# round 2, call O: full-pipeline workload (BASELINE configs[4] per-GPU share), bench stdout check
mkdir -p gpurun_out
timeout 900 python bench.py --workload pipeline --steps 2 --warmup 1 > gpurun_out/r02_bench_pipeline_1gpu.json 2> gpurun_out/r02_bench_pipeline_1gpu.err; echo "rc=$?"; wc -l gpurun_out/r02_bench_pipeline_1gpu.json; tail -c 900 gpurun_out/r02_bench_pipeline_1gpu.json; tail -3 gpurun_out/r02_bench_pipeline_1gpu.err | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 3 2> gpurun_out/r2o_bench.err | wc -l
