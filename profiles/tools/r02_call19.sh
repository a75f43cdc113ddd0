#!/bin/bash
# meshes: unfused generation loop (k_shade leaves children untraced, k_intersect per generation) against the fused
# trace-ahead (librpx_prev.so / RPX_MESH_FUSED=1); k_intersect<MESH> at 4 / 6 / 8 CTAs per SM; full GPU suite
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_c19_tests.log 2>&1
{
L="librpx_prev.so librpx.so librpx_i6.so librpx_i8.so"
bash profiles/tools/ab1.sh "$L $L" "mesh"
bash profiles/tools/ab1.sh "$L" "mesh_large"
RPX_MESH_FUSED=1 bash profiles/tools/ab1.sh "librpx.so" "mesh" | sed "s/librpx.so/librpx.so[RPX_MESH_FUSED=1]/"
} > gpurun_out/r02_c19_ab.log 2>&1
cat gpurun_out/r02_c19_tests.log gpurun_out/r02_c19_ab.log
