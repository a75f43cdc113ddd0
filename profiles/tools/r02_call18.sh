#!/bin/bash
# meshes: while-while walk (librpx.so) against the first packed walk (librpx_prev.so), single-triangle leaves (new default)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
L="librpx_prev.so librpx.so"
bash profiles/tools/ab1.sh "$L $L" "mesh"
bash profiles/tools/ab1.sh "$L" "mesh_large"
RPX_BVH_LEAF=2 bash profiles/tools/ab1.sh "librpx.so" "mesh" | sed "s/librpx.so/leaf=2/"
RPX_BVH_LEAF=4 bash profiles/tools/ab1.sh "librpx.so" "mesh" | sed "s/librpx.so/leaf=4/"
timeout 600 python -m pytest tests -m gpu -x -q -k "mesh or uvpatch or golden" 2>&1 | tail -3
} > gpurun_out/r02_c18_ab.log 2>&1
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    -k regex:"k_shade|k_intersect" -c 3 --csv --log-file gpurun_out/r02_c18_div.csv python bench.py --workload mesh --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
cat gpurun_out/r02_c18_ab.log; grep -v "^==" gpurun_out/r02_c18_div.csv | cut -d, -f5,13,15 | cut -c1-200
