#!/bin/bash
# two-rank run of the bench (one process per GPU, NCCL all-gather of the result), plus the reference arm contract under torchrun
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/m2_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err; tail -c 700 gpurun_out/m2_bench.json; tail -5 gpurun_out/m2_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/m2_ref.json 2> gpurun_out/m2_ref.err; tail -c 400 gpurun_out/m2_ref.json
echo done
