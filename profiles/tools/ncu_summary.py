"""Condense `ncu -i X.ncu-rep --page raw --csv` into the one-row-per-launch summary kept under profiles/.

usage: ncu -i X.ncu-rep --page raw --csv | python profiles/tools/ncu_summary.py > profiles/rNN_ncu_summary.csv
"""
import csv
import re
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
STALL = re.compile(r"^smsp__pcsamp_warps_issue_stalled_([a-z_]+)$")

rows = list(csv.reader(sys.stdin))
head, units, data = rows[0], rows[1], rows[2:]
idx = {name: i for i, name in enumerate(head)}
cols = [("Kernel Name", idx["Kernel Name"])] + [(k, idx[k]) for k in KEEP if k in idx]
cols += sorted((("samples_" + m.group(1)), i) for i, name in enumerate(head) for m in [STALL.match(name)] if m)
w = csv.writer(sys.stdout)
w.writerow([c for c, _ in cols])
w.writerow([units[i] if c != "Kernel Name" else "" for c, i in cols])
for r in data:
    out = []
    for c, i in cols:
        v = r[i]
        if c == "Kernel Name":
            v = re.sub(r"rpx::", "", v)
            v = re.sub(r"\((bool|int|unsigned int)\)", "", v)
        out.append(v)
    w.writerow(out)
