#!/bin/bash
# Gausslets: second parabasal loop reads the hit distances / child constants back from where they already sit
# (RPX_G_RELOAD, librpx.so) against the build before it (librpx_base.so), + the software-pipelined first loop
# (librpx_pipe.so).  Plain rays: lean child staging at 4 CTAs / SM (librpx_lean4.so) against the shipped staging
# on every plain-ray workload.  Parity subsets under the candidate libraries; drop-in timing of the default workload.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
ab() {  # ab "<libs>" "<workloads>" tag
  for w in $2; do for l in $1; do
    RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
        > $O/r02_c21_ab_${w}_${l%.so}.log 2>&1
    tail -1 $O/r02_c21_ab_${w}_${l%.so}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], '%.3f'%d['roofline']['frac'])" || echo "$w $l FAILED"
  done; done
}
{
ab "librpx_base.so librpx.so librpx_pipe.so" "config5_1e6"
ab "librpx_base.so librpx_lean4.so" "config4_prisms config5_rays config4_grating config3 mesh"
for l in librpx.so librpx_pipe.so; do
  echo "parity under $l"
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_properties_gpu.py tests/test_parity_fullsize_gpu.py -m gpu -x -q -k "config5 or zoo or big_scene or mesh or uvpatch or streaming or baseline_size" 2>&1 | tail -2
done
echo "parity under librpx_lean4.so"
RPX_LIB=$PWD/raypier_optics_b200/csrc/librpx_lean4.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
} > $O/r02_c21_ab.log 2>&1
(time timeout 600 python bench.py --rays 8000000 --steps 3 --warmup 3 --no-cpu-baseline) > $O/r02_c21_bench_consume8e6.log 2>&1
cat $O/r02_c21_ab.log
grep "^{" $O/r02_c21_bench_consume8e6.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e_dropin'])"
