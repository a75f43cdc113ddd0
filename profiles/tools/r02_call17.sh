#!/bin/bash
# meshes: leaf-size sweep of the host BVH, the STL-sized scene, ncu --set full (with source) of the mesh kernels
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
for leaf in 1 2 4 8; do
    RPX_BVH_LEAF=$leaf bash profiles/tools/ab1.sh "librpx.so" "mesh" | sed "s/librpx.so/leaf=$leaf/"
done
bash profiles/tools/ab1.sh "librpx.so" "mesh_large"
RPX_BVH_LEAF=2 bash profiles/tools/ab1.sh "librpx.so" "mesh_large" | sed "s/librpx.so/leaf=2/"
RPX_MESH_F64=1 bash profiles/tools/ab1.sh "librpx.so" "mesh_large" | sed 's/librpx.so/librpx.so[RPX_MESH_F64=1]/'
} > gpurun_out/r02_c17_ab.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_shade|k_intersect" -c 3 -f -o gpurun_out/prof_r02_mesh \
    python bench.py --workload mesh --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_c17_ncu.log 2>&1
cat gpurun_out/r02_c17_ab.log; tail -2 gpurun_out/r02_c17_ncu.log
