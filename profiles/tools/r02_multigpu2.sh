#!/bin/bash
# 2-GPU validation of the final round-2 build (run with gpurun --gpus 2): on-GPU self-check of the distributed paths,
# then the default bench command at N=2 exactly as the driver launches it (full 1.25e8 gausslets per GPU, short).
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
(nvidia-smi -L; free -g | head -2; nproc) > gpurun_out/r02_mg2_box.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > gpurun_out/r02_mg2_check.log 2>&1
echo "multi_gpu_check rc=$?" >> gpurun_out/r02_mg2_check.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 3 --warmup 3) > gpurun_out/r02_mg2_bench2.log 2>&1
cat gpurun_out/r02_mg2_box.txt; tail -4 gpurun_out/r02_mg2_check.log; tail -c 1500 gpurun_out/r02_mg2_bench2.log
