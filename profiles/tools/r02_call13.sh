#!/bin/bash
# gausslet k_shade experiments: L2 prefetch of the parabasal rows (e1), 16-byte child stores (e2), unconditional
# parabasal loads (e4), parabasal intersections before the material (e5), combinations
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L="librpx.so librpx_e1.so librpx_e2.so librpx_e5.so librpx_e1245.so"
bash profiles/tools/ab1.sh "$L $L" "config5_1e6" > gpurun_out/r02_c13_ab.log 2>&1
cat gpurun_out/r02_c13_ab.log
