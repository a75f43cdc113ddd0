#!/bin/bash
# new default build (parabasal-first gausslet order, facet record side array, packed fp32 BVH) against the previous
# commit's build (librpx_head.so): full GPU test suite, mesh A/B (also RPX_MESH_F64=1 = side array only), regression A/B
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_c16_tests.log 2>&1
L="librpx_head.so librpx.so"
bash profiles/tools/ab1.sh "$L $L" "mesh" > gpurun_out/r02_c16_ab.log 2>&1
RPX_MESH_F64=1 bash profiles/tools/ab1.sh "librpx.so" "mesh" | sed 's/librpx.so/librpx.so[RPX_MESH_F64=1]/' >> gpurun_out/r02_c16_ab.log 2>&1
bash profiles/tools/ab1.sh "$L $L" "config2 config5_1e6 config4_prisms config3" >> gpurun_out/r02_c16_ab.log 2>&1
cat gpurun_out/r02_c16_tests.log gpurun_out/r02_c16_ab.log
