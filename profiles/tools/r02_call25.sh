#!/bin/bash
# Gausslet k_shade with the two parabasal loops ROLLED (librpx.so, 103 KB of SASS instead of 205 KB) and unrolled by 2
# (librpx_u2.so), both with the hoisted Snell ratios, against the fully unrolled base (librpx_base.so); ncu stall
# summary of the rolled build; gausslet parity under both.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
{
for w in config5_1e6; do for l in librpx_base.so librpx.so librpx_u2.so; do
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
      > $O/r02_c25_ab_${w}_${l%.so}.log 2>&1
  tail -1 $O/r02_c25_ab_${w}_${l%.so}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], '%.3f'%d['roofline']['frac'])" || echo "$w $l FAILED"
done; done
for l in librpx.so librpx_u2.so; do
  echo "parity under $l"
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_properties_gpu.py tests/test_parity_fullsize_gpu.py tests/test_sequence.py tests/test_consume_gpu.py -m gpu -x -q -k "config5 or zoo or big_scene or mesh or uvpatch or streaming or baseline_size or gausslet or decomposition or consume" 2>&1 | tail -2
done
} > $O/r02_c25_ab.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 3 -f -o $O/prof_r02d_gauss_rolled \
    python bench.py --workload config5_1e6 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02_c25_ncu.log 2>&1
cat $O/r02_c25_ab.log
