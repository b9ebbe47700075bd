#!/bin/bash
# secondary measurement: the bare denoiser without CFG ("no-CFG variant")
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 90 python bench.py --no-cfg --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nocfg_bench.json 2> gpurun_out/nocfg_bench.err; tail -c 400 gpurun_out/nocfg_bench.json; tail -3 gpurun_out/nocfg_bench.err
echo done
