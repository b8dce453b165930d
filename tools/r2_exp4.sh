#!/bin/bash
O=gpurun_out/exp4; mkdir -p $O
for p in fused single single_notma three_pass; do
  echo "== $p" | tee -a $O/check.log
  QLB_PIPELINE=$p timeout 300 python tools/gpu_check.py --config C3 --batch 32768 >> $O/check.log 2>&1; echo "rc=$?" >> $O/check.log
done
for p in fused single; do
echo "== $p C5" >> $O/check.log
QLB_PIPELINE=$p timeout 300 python tools/gpu_check.py --config C5 --batch 32768 --time-batch 2097152 >> $O/check.log 2>&1
echo "== $p C2" >> $O/check.log
QLB_PIPELINE=$p timeout 300 python tools/gpu_check.py --config C2 --batch 32768 --time-batch 65536 >> $O/check.log 2>&1
done
grep -E "==|rc=|device-resident|flag mism|grf rel|Error|error" $O/check.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
for p in fused single; do
QLB_PIPELINE=$p timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qlb_ -s 12 -c 6 --csv --log-file $O/launches_$p.csv python tools/gpu_check.py --config C3 --batch 1024 > /dev/null 2>&1
grep -E "qlb_" $O/launches_$p.csv | awk -F'","' '{print $5, $NF}' | tail -6
QLB_PIPELINE=$p timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlb_ -s 12 -c 3 -o $O/prof_$p python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu_$p.log 2>&1
done
ls -la $O
