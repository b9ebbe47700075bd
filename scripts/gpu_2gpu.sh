#!/bin/bash
# two-rank run of the bench as the driver launches it (one process per GPU, NCCL all-gather of the result; weak value + strong config3)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/m2_smi.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err; tail -c 900 gpurun_out/m2_bench.json; tail -3 gpurun_out/m2_bench.err | cut -c1-200
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/m2_ref.json 2> gpurun_out/m2_ref.err; tail -c 400 gpurun_out/m2_ref.json
echo done
