#!/bin/bash
# ncu --set full with source of the config3 kernels (aspheric Newton + Zernike secant; k_shade<0,1,...> is 499 KB of
# SASS): are instruction-cache misses a factor there too?
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_shade|k_intersect" -c 3 -f -o gpurun_out/prof_r02e_config3 \
    python bench.py --workload config3 --rays 1000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_c27_ncu.log 2>&1
ls -la gpurun_out/prof_r02e*
