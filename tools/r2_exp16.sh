#!/bin/bash
# experiment: what would lock-stepped warp pairs (shared instruction fetch on a sub-partition) buy?  vT = both warps of a
# pair walk through the same boxes (twice the work, natural lock step), vS = the same static distribution, distinct work.
O=gpurun_out/exp16; mkdir -p $O
bash tools/r2_variants.sh exp16 "vS vT main vS vT"
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
for v in vS vT; do
  QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_$v.so timeout 300 ncu --metrics $M --clock-control none -k regex:qlb_single -s 3 -c 1 --csv --log-file $O/ncu_$v.csv python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu_$v.log 2>&1
  echo "== $v"; grep -E "no_instruction|wait_per|issue_active.avg|time_duration|fp64|inst_executed" $O/ncu_$v.csv | awk -F'","' '{print $(NF-2), $NF}'
done
