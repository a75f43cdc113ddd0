#!/bin/bash
# Gausslet k_shade: parabasal origins / directions of the tile fetched with cp.async into the (idle) child staging
# buffer at tile start (librpx_ps.so) against the shipped build; gausslet parity under it.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
{
for w in config5_1e6; do for l in librpx.so librpx_ps.so; do
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
      > $O/r02_c30_ab_${w}_${l%.so}.log 2>&1
  tail -1 $O/r02_c30_ab_${w}_${l%.so}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], '%.3f'%d['roofline']['frac'])" || echo "$w $l FAILED"
done; done
echo "parity under librpx_ps.so"
RPX_LIB=$PWD/raypier_optics_b200/csrc/librpx_ps.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_sequence.py tests/test_parity_fullsize_gpu.py tests/test_consume_gpu.py -m gpu -x -q -k "config5 or zoo or big_scene or gausslet or decomposition or consume or baseline_size" 2>&1 | tail -2
} > $O/r02_c30_ab.log 2>&1
cat $O/r02_c30_ab.log
