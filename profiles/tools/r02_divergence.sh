#!/bin/bash
# Round-2 divergence / stall evidence for the mixed scenes: per kernel launch, the average
# number of active threads per executed warp instruction, the issue-slot and fp64-pipe utilisation and the
# main stall reasons, for prisms, Michelson rays and Michelson gausslets.  Run on a B200 through gpurun.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
for w in config4_prisms config5_rays config5_1e6 config2; do
    timeout 600 ncu --clock-control none --metrics $M -k regex:"k_shade" -c 12 --csv --log-file gpurun_out/r02_div_$w.csv \
        python bench.py --workload $w --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_div_$w.log 2>&1
    tail -1 gpurun_out/r02_div_$w.log | cut -c1-200
done
