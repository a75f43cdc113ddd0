#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CSRC=raypier_optics_b200/csrc
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_c8_tests.log 2>&1
{
for l in librpx_te.so librpx_mb3.so librpx_te3.so; do
    [ -f $CSRC/$l ] || continue
    RPX_LIB=$PWD/$CSRC/$l timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_properties_gpu.py \
        -m gpu -x -q -k "config5 or zoo or big_scene or streaming or uvpatch" 2>&1 | tail -2
done
libs=""; for l in librpx.so librpx_te.so librpx_mb3.so librpx_te3.so; do [ -f $CSRC/$l ] && libs="$libs $l"; done
bash profiles/tools/ab1.sh "$libs $libs" "config5_1e6"
bash profiles/tools/ab1.sh "librpx.so" "config2 config4_prisms config5_rays config4_grating config3 config1"
} > gpurun_out/r02_c8_ab.log 2>&1
(time timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline) > gpurun_out/r02_c8_bench_full.log 2>&1
tail -8 gpurun_out/r02_c8_tests.log; cat gpurun_out/r02_c8_ab.log; tail -c 400 gpurun_out/r02_c8_bench_full.log
