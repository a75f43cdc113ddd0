#!/bin/bash
# Why does the gausslet k_shade lose 16 % when work is REMOVED from its second parabasal loop?  Per-launch time,
# executed instructions and warp-state statistics of the base build and of the hoisted-Snell-ratio build.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
for l in librpx_base.so librpx.so; do
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 300 ncu --clock-control none --section WarpStateStats --section SchedulerStats \
      --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio \
      -k regex:k_shade -c 8 --csv --log-file $O/r02_c23_${l%.so}.csv python bench.py --workload config5_1e6 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02_c23_${l%.so}.log 2>&1
done
ls -la $O/r02_c23*
