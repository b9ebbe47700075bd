#!/bin/bash
# final default bench line (with the per-launch roofline at the configs[2] launch shapes)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_default_line.json 2> gpurun_out/r02_bench_default_line.err; tail -c 600 gpurun_out/r02_bench_default_line.json; tail -2 gpurun_out/r02_bench_default_line.err | cut -c1-200
