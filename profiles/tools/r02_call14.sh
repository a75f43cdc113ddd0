#!/bin/bash
# second batch of gausslet k_shade variants (parabasal intersections before the material = e5, base-ray fields loaded
# after them = e6, each with unconditional parabasal loads = e45 / e46), parity of the candidates, and an ncu
# --set full capture (with source) of the plain-ray k_shade on the achromat
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CSRC=raypier_optics_b200/csrc
L="librpx.so librpx_e5.so librpx_e6.so librpx_e45.so librpx_e46.so"
bash profiles/tools/ab1.sh "$L $L" "config5_1e6" > gpurun_out/r02_c14_ab.log 2>&1
{
for l in librpx_e5.so librpx_e6.so; do
    RPX_LIB=$PWD/$CSRC/$l timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_properties_gpu.py \
        -m gpu -x -q -k "config5 or zoo or big_scene or streaming or uvpatch or mesh or gauss" 2>&1 | tail -2
done
} > gpurun_out/r02_c14_parity.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -s 1 -c 2 -f -o gpurun_out/prof_r02_config2 \
    python bench.py --workload config2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_c14_ncu.log 2>&1
cat gpurun_out/r02_c14_ab.log gpurun_out/r02_c14_parity.log; tail -3 gpurun_out/r02_c14_ncu.log
