#!/bin/bash
# Runs ON THE GPU BOX: a handful of ncu metrics of one kernel family on the 32-image driver (one launch).
#   tools/ncu_quick.sh TAG KERNEL_REGEX      -> gpurun_out/TAG_quick.csv
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,\
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,\
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,\
l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,\
l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__lsu_writeback_active_mem_lgds.sum,\
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,\
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,\
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,\
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__sass_inst_executed_op_shared_ld.sum,smsp__sass_inst_executed_op_shared_st.sum,\
smsp__inst_executed_op_shared_atom.sum,smsp__sass_inst_executed_op_global_ld.sum
ncu --metrics $M --clock-control none -k regex:"$2" -c 1 --csv --log-file gpurun_out/$1_quick.csv python tools/prof_driver.py ${3:-32} $4 $5 > /dev/null 2>&1
python - <<P
import csv
rows = [r for r in csv.reader(open("gpurun_out/$1_quick.csv")) if len(r) > 10]
h = rows[0]; i_n, i_v = h.index("Metric Name"), h.index("Metric Value")
for r in rows[1:]:
    print(f"{r[i_n][:88]:88s} {r[i_v]}")
P
