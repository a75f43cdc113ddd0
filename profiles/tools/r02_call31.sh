#!/bin/bash
# Meshes: 4-wide packed BVH nodes (librpx.so) against the 2-wide nodes of the round-2 walk so far (librpx_bvh2.so);
# mesh / UV patch parity and golden tests under the new walk, then the full GPU suite.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "mesh or uvpatch" 2>&1 | tail -3) > $O/r02_c31_parity.log 2>&1
{
for w in mesh mesh_large; do for l in librpx_bvh2.so librpx.so; do
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
      > $O/r02_c31_ab_${w}_${l%.so}.log 2>&1
  tail -1 $O/r02_c31_ab_${w}_${l%.so}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], d['trace']['generations'])" || echo "$w $l FAILED"
done; done
} > $O/r02_c31_ab.log 2>&1
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > $O/r02_c31_tests.log 2>&1
cat $O/r02_c31_parity.log $O/r02_c31_ab.log $O/r02_c31_tests.log
