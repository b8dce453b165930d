#!/bin/bash
# round-2 experiment 3: the two-kernel pipeline (tile kernel with TMA staging + rounds kernel fed from the stash)
O=gpurun_out/exp3; mkdir -p $O
for p in fused fused_notma three_pass; do
  echo "== $p" | tee -a $O/check.log
  QLB_PIPELINE=$p timeout 300 python tools/gpu_check.py --config C3 --batch 32768 >> $O/check.log 2>&1; echo "rc=$?" >> $O/check.log
done
echo "== fused C5" >> $O/check.log
timeout 300 python tools/gpu_check.py --config C5 --batch 32768 --time-batch 2097152 >> $O/check.log 2>&1
echo "== fused C2" >> $O/check.log
timeout 300 python tools/gpu_check.py --config C2 --batch 32768 --time-batch 65536 >> $O/check.log 2>&1
echo "== rounds2" >> $O/check.log
QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_rounds2.so timeout 300 python tools/gpu_check.py --config C3 --batch 32768 >> $O/check.log 2>&1
grep -E "==|rc=|device-resident|status hist|flag mism|grf rel|Error|error" $O/check.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qlb_ -s 12 -c 9 --csv --log-file $O/launches.csv python tools/gpu_check.py --config C3 --batch 1024 > /dev/null 2>&1
grep -E "qlb_" $O/launches.csv | awk -F'","' '{print $5, $NF}' | tail -9
QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_rounds2.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qlb_ -s 12 -c 3 --csv --log-file $O/launches_rounds2.csv python tools/gpu_check.py --config C3 --batch 1024 > /dev/null 2>&1
grep -E "qlb_" $O/launches_rounds2.csv | awk -F'","' '{print $5, $NF}' | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlb_ -s 12 -c 3 -o $O/prof python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu.log 2>&1
python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; head -c 400 $O/bench.json
ls -la $O
