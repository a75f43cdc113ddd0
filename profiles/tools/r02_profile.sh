#!/bin/bash
# Round-2 profiling pass (run on a B200 through gpurun): A/B of the gausslet k_shade variants, ncu launch
# list of the streamed north-star step, ncu --set full of the gausslet kernel, fp64 instruction counts of the
# config3 kernels.  Outputs land in gpurun_out/ (copied into profiles/ by hand afterwards).
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CSRC=raypier_optics_b200/csrc
{
echo "== parity of the gausslet cases under the variant libraries"
for l in librpx_cs.so librpx_csb.so librpx_b2.so; do
    [ -f $CSRC/$l ] || continue
    RPX_LIB=$PWD/$CSRC/$l timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_properties_gpu.py \
        -m gpu -x -q -k "config5 or zoo or big_scene or streaming or uvpatch" 2>&1 | tail -2
done
echo "== A/B (Michelson gausslets, 1e6 per generation-0)"
libs=""; for l in librpx.so librpx_cs.so librpx_csb.so librpx_b2.so; do [ -f $CSRC/$l ] && libs="$libs $l"; done
bash profiles/tools/ab1.sh "$libs $libs" "config5_1e6"
} > gpurun_out/r02_ab_gauss.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --rays 4000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 3 -f -o gpurun_out/prof_r02_gauss \
    python bench.py --workload config5_1e6 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_gauss.log 2>&1
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    -k regex:"k_shade|k_intersect" -c 4 --csv --log-file gpurun_out/r02_fp64_config3.csv \
    python bench.py --workload config3 --rays 1000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_c3.log 2>&1
cat gpurun_out/r02_ab_gauss.log
tail -3 gpurun_out/r02_ncu_gauss.log gpurun_out/r02_ncu_c3.log
