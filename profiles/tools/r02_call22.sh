#!/bin/bash
# New default (Snell ratios of the parabasal children hoisted out of the six-ray loop; lean child staging for plain
# rays only) against the build before both (librpx_base.so); full GPU suite on the new default.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
{
for w in config5_1e6 config2 config4_prisms; do for l in librpx_base.so librpx.so; do
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
      > $O/r02_c22_ab_${w}_${l%.so}.log 2>&1
  tail -1 $O/r02_c22_ab_${w}_${l%.so}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], '%.3f'%d['roofline']['frac'])" || echo "$w $l FAILED"
done; done
} > $O/r02_c22_ab.log 2>&1
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/r02_c22_tests.log 2>&1
cat $O/r02_c22_ab.log $O/r02_c22_tests.log
