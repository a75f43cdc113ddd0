#!/bin/bash
# The experiments that round 1 prepared but could not run (its GPU minutes were spent).
#
#   1. HERE (no GPU needed):   profiles/tools/prepared_experiments.sh build
#   2. on a B200:              gpurun --timeout 600 -- 'bash profiles/tools/prepared_experiments.sh run'
#   3. read gpurun_out/prepared.log
#
# build: librpx_lean.so = -DRPX_LEAN_STAGE=1 -DRPX_MIN_BLOCKS=5 (lean child staging, 5 CTAs / SM)
#        librpx_mo.so   = -DRPX_MESH_ORDERED=1 (near-child-first BVH walk for triangle meshes)
#        librpx_t64.so  = -DRPX_TILE=64 -DRPX_MIN_BLOCKS=8 (64-ray tiles, 8 CTAs / SM)
# run:   (a) parity + golden suites under the lean library, (b) A/B of the two libraries on config2 /
#        prisms / Michelson gausslets, (c) the prepared GPU tests (in-place generation 0 of
#        rpx_trace_streamed), (d) the e2e arm with and without --e2e-inplace.
set -u
cd "$(dirname "$0")/../.."
CSRC=raypier_optics_b200/csrc
case "${1:-}" in
build)
    make -C $CSRC -j"$(nproc)" OBJDIR=obj_lean LIB=librpx_lean.so RPX_EXTRA="-DRPX_LEAN_STAGE=1 -DRPX_MIN_BLOCKS=5" \
        2>&1 | grep -iE "error|warning" ; ls -la $CSRC/librpx_lean.so
    make -C $CSRC -j"$(nproc)" OBJDIR=obj_t64 LIB=librpx_t64.so RPX_EXTRA="-DRPX_TILE=64 -DRPX_MIN_BLOCKS=8" \
        2>&1 | grep -iE "error|warning" ; ls -la $CSRC/librpx_t64.so
    make -C $CSRC -j"$(nproc)" OBJDIR=obj_mo LIB=librpx_mo.so RPX_EXTRA="-DRPX_MESH_ORDERED=1" \
        2>&1 | grep -iE "error|warning" ; ls -la $CSRC/librpx_mo.so
    ;;
run)
    mkdir -p gpurun_out
    {
        echo "== (a) parity under the lean library"
        RPX_LIB=$PWD/$CSRC/librpx_lean.so timeout 200 python -m pytest tests/test_parity_gpu.py tests/test_golden.py \
            tests/test_properties_gpu.py -m gpu -x -q 2>&1 | tail -3
        echo "== (b) A/B default vs lean"
        bash profiles/tools/ab1.sh "librpx.so librpx_lean.so librpx_t64.so librpx.so librpx_lean.so librpx_t64.so" "config2"
        bash profiles/tools/ab1.sh "librpx.so librpx_lean.so librpx_t64.so" "config4_prisms config5"
        echo "== (a2) parity under the 64-ray-tile library"
        RPX_LIB=$PWD/$CSRC/librpx_t64.so timeout 200 python -m pytest tests/test_parity_gpu.py tests/test_golden.py \
            -m gpu -x -q 2>&1 | tail -3
        echo "== (b2) ordered mesh walk: parity of the mesh cases, then A/B on the 71k-facet scene"
        RPX_LIB=$PWD/$CSRC/librpx_mo.so timeout 200 python -m pytest tests/test_parity_gpu.py tests/test_golden.py \
            -m gpu -x -q -k "mesh" 2>&1 | tail -3
        bash profiles/tools/ab1.sh "librpx.so librpx_mo.so" "mesh"
        echo "== (c) prepared GPU tests"
        timeout 200 python -m pytest tests -m gpu_prepared -x -q 2>&1 | tail -3
        echo "== (d) e2e, separate buffers vs in place"
        for flag in "" "--e2e-inplace"; do
            timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $flag 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('e2e $flag', '%.4g' % e['value'], 'd2h bytes', e['d2h_bytes_per_step'])"
        done
    } 2>&1 | tee gpurun_out/prepared.log
    ;;
*)
    sed -n 2,14p "$0"
    ;;
esac
