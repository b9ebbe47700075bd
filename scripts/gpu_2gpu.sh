#!/bin/bash
# two-rank run of the bench (one process per GPU, NCCL all-gather of the result)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/m2_smi.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err; tail -c 700 gpurun_out/m2_bench.json; tail -5 gpurun_out/m2_bench.err
echo done
