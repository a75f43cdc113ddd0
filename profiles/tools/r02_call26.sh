#!/bin/bash
# Gausslet k_shade: rolled + software-pipelined parabasal loops (p1: unroll 1, p2: unroll 2) and unroll 3 against
# unroll 2 (best so far) and the fully unrolled base.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
{
for w in config5_1e6; do for l in librpx_base.so librpx_u2.so librpx_p1.so librpx_p2.so librpx_u3.so; do
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
      > $O/r02_c26_ab_${w}_${l%.so}.log 2>&1
  tail -1 $O/r02_c26_ab_${w}_${l%.so}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], '%.3f'%d['roofline']['frac'])" || echo "$w $l FAILED"
done; done
for l in librpx_p1.so librpx_p2.so; do
  echo "parity under $l"
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_sequence.py -m gpu -x -q -k "config5 or zoo or big_scene or mesh or uvpatch or gausslet or decomposition" 2>&1 | tail -2
done
} > $O/r02_c26_ab.log 2>&1
cat $O/r02_c26_ab.log
