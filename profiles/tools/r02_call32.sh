#!/bin/bash
# Gausslet k_shade on the ROLLED kernel: three variants whose earlier (negative) measurements were taken on the
# layout-sensitive fully unrolled kernel -- L2-only parent reads, lean child staging, 3 CTAs / SM at 168 registers.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
{
for w in config5_1e6; do for l in librpx.so librpx_gldcg.so librpx_glean.so librpx_gb3.so; do
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
      > $O/r02_c32_ab_${w}_${l%.so}.log 2>&1
  tail -1 $O/r02_c32_ab_${w}_${l%.so}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], '%.3f'%d['roofline']['frac'])" || echo "$w $l FAILED"
done; done
} > $O/r02_c32_ab.log 2>&1
cat $O/r02_c32_ab.log
