#!/bin/bash
# round-2 experiment 2: first run of the fused kernel
mkdir -p gpurun_out/exp2
O=gpurun_out/exp2
for p in fused_notma fused three_pass; do
  echo "== $p" | tee -a $O/check.log
  QLB_PIPELINE=$p timeout 300 python tools/gpu_check.py --config C3 --batch 32768 >> $O/check.log 2>&1; echo "rc=$?" >> $O/check.log
done
QLB_PIPELINE=fused timeout 300 python tools/gpu_check.py --config C5 --batch 32768 --time-batch 2097152 >> $O/check.log 2>&1
QLB_PIPELINE=fused timeout 300 python tools/gpu_check.py --config C2 --batch 32768 --time-batch 65536 >> $O/check.log 2>&1
grep -E "==|rc=|device-resident|status hist|flag mism|grf rel|Error|error" $O/check.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
QLB_PIPELINE=fused timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qlb_ -s 8 -c 6 --csv --log-file $O/launches_fused.csv python tools/gpu_check.py --config C3 --batch 1024 > /dev/null 2>&1
grep -E "qlb_" $O/launches_fused.csv | awk -F'","' '{print $5, $NF}' | tail -6
QLB_PIPELINE=fused timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlb_fused -s 4 -c 1 -o $O/prof_fused python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu.log 2>&1
ls -la $O
