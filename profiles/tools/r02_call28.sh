#!/bin/bash
# Gausslet k_shade: unroll factors of the first (intersections) and second (children) parabasal loop chosen separately
# (m12 = 1 and 2, ...), against the shipped 2 / 2; then ncu --set full of the shipped gausslet and plain-ray kernels.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
{
for w in config5_1e6; do for l in librpx.so librpx_m12.so librpx_m21.so librpx_m32.so librpx_m23.so librpx.so; do
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
      > $O/r02_c28_ab_${w}_${l%.so}.log 2>&1
  tail -1 $O/r02_c28_ab_${w}_${l%.so}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], '%.3f'%d['roofline']['frac'])" || echo "$w $l FAILED"
done; done
} > $O/r02_c28_ab.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 3 -f -o $O/prof_r02f_gauss_shipped \
    python bench.py --workload config5_1e6 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02_c28_ncu_gauss.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 2 -f -o $O/prof_r02f_config2_shipped \
    python bench.py --workload config2 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02_c28_ncu_config2.log 2>&1
cat $O/r02_c28_ab.log
