#!/bin/bash
# 2-GPU validation of the distributed paths (run with gpurun --gpus 2)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_mg_gpus.txt 2>&1
(free -g | head -2; nproc) >> gpurun_out/r02_mg_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > gpurun_out/r02_mg_check.log 2>&1
echo "multi_gpu_check rc=$?" >> gpurun_out/r02_mg_check.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 3 --warmup 1 --rays 40000000) > gpurun_out/r02_mg_bench2.log 2>&1
(time timeout 600 python bench.py --gpus 1 --steps 3 --warmup 1 --rays 40000000 --no-cpu-baseline) > gpurun_out/r02_mg_bench1.log 2>&1
tail -4 gpurun_out/r02_mg_check.log; tail -c 600 gpurun_out/r02_mg_bench2.log
