#!/bin/bash
# third batch: look-back by the first warp to finish its trace-ahead (lb; plain rays and gausslets), software-pipelined
# parabasal loops on top of the parabasal-first order (e5p1 = second loop, e5p2 = first loop)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CSRC=raypier_optics_b200/csrc
L="librpx.so librpx_e5.so librpx_e5p1.so librpx_e5p2.so librpx_lb.so"
bash profiles/tools/ab1.sh "$L $L" "config5_1e6" > gpurun_out/r02_c15_ab.log 2>&1
L="librpx.so librpx_lb.so"
bash profiles/tools/ab1.sh "$L $L" "config2 config4_prisms config5_rays" >> gpurun_out/r02_c15_ab.log 2>&1
{
for l in librpx_lb.so; do
    RPX_LIB=$PWD/$CSRC/$l timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_properties_gpu.py \
        -m gpu -x -q 2>&1 | tail -2
done
} > gpurun_out/r02_c15_parity.log 2>&1
cat gpurun_out/r02_c15_ab.log gpurun_out/r02_c15_parity.log
