#!/bin/bash
# diagnostic: is the instruction-fetch stall caused by two warps of one SM sub-partition running different halves of
# the code?  One CTA (four warps, one per sub-partition) per SM against the default two.
O=gpurun_out/exp13; mkdir -p $O
echo "== default" >> $O/check.log
timeout 300 python tools/gpu_check.py --config C3 --batch 8192 >> $O/check.log 2>&1
echo "== one CTA per SM" >> $O/check.log
QLB_FUSED_BPS=1 timeout 300 python tools/gpu_check.py --config C3 --batch 8192 >> $O/check.log 2>&1
grep -E "==|device-resident|flag mism" $O/check.log
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
for b in 2 1; do
  QLB_FUSED_BPS=$b timeout 300 ncu --metrics $M --clock-control none -k regex:qlb_single -s 3 -c 1 --csv --log-file $O/ncu_bps$b.csv python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu_bps$b.log 2>&1
  echo "== bps $b"; grep -E "no_instruction|wait_per|issue_active.avg|time_duration|fp64" $O/ncu_bps$b.csv | awk -F'","' '{print $(NF-2), $NF}'
done
