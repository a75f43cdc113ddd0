#!/bin/bash
# Meshes: binned-SAH split in the host BVH builder (RPX_BVH_SAH=1) against the median split; parity of the mesh / UV
# patch cases under it.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
{
RPX_BVH_SAH=1 timeout 60 python -m pytest tests/test_parity_gpu.py tests/test_sequence.py -m gpu -x -q -k "mesh or uvpatch" 2>&1 | tail -2
for v in 0 1; do
  RPX_BVH_SAH=$v timeout 60 python bench.py --workload mesh --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_c35_mesh_sah$v.log 2>&1
  tail -1 $O/r02_c35_mesh_sah$v.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mesh sah=$v', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], d['trace']['generations'])" || echo "mesh $v FAILED"
done
RPX_BVH_SAH=1 timeout 80 python bench.py --workload mesh_large --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_c35_mesh_large_sah1.log 2>&1
tail -1 $O/r02_c35_mesh_large_sah1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mesh_large sah=1', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], d['trace']['generations'])" || echo "mesh_large FAILED"
} > $O/r02_c35.log 2>&1
cat $O/r02_c35.log
