#!/bin/bash
# a few issue/divergence metrics of the first trace launches under each traversal variant
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,smsp__average_warp_latency_per_inst_issued.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
for cfg in "$@"; do
  set -- $cfg
  SPT_TRACE_VARIANT=$1 SPT_LEAF_WAIT=$2 ncu --metrics $M --clock-control none -k regex:k_trace -s 48 -c 6 --csv --log-file gpurun_out/ncu_var_$1_$2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
done
