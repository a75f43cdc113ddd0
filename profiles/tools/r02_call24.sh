#!/bin/bash
# ncu --set full with source of the hoisted-Snell-ratio build (first two k_shade launches of the 1e6-gausslet trace)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 2 -f -o gpurun_out/prof_r02c_gauss_hoist \
    python bench.py --workload config5_1e6 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_c24_ncu.log 2>&1
ls -la gpurun_out/prof_r02c*
