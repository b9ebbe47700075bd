#!/bin/bash
# round 2, call H (2 GPUs): the bench line at N = 2 as the driver launches it (weak scaling + config3 strong scaling), reference arm at N = 2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err
echo "rc=$?"; tail -c 2000 gpurun_out/r2h_bench_2gpu.json; tail -5 gpurun_out/r2h_bench_2gpu.err | cut -c1-300
