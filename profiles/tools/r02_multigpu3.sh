#!/bin/bash
# 2-GPU self-check of the distributed paths after the capacity fix of its timed terminal gather (gpurun --gpus 2)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > gpurun_out/r02_mg3_check.log 2>&1
echo "multi_gpu_check rc=$?" >> gpurun_out/r02_mg3_check.log
grep "multi_gpu_check" gpurun_out/r02_mg3_check.log
