#!/bin/bash
# variants of the fused kernel: stash capacity / hysteresis, CTAs per SM; parity + timing on C3, C5, C2 and the FP32 twin
O=gpurun_out/exp10; mkdir -p $O
for v in main cap15 low12 c3 t256; do
  if [ $v = main ]; then export QLB_LIB=$PWD/quadruped_locomotion_b200/libqlb.so; else export QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_$v.so; fi
  echo "== $v C3" >> $O/check.log
  timeout 300 python tools/gpu_check.py --config C3 --batch 32768 >> $O/check.log 2>&1; echo "rc=$?" >> $O/check.log
  echo "== $v C5" >> $O/check.log
  timeout 300 python tools/gpu_check.py --config C5 --batch 32768 --time-batch 2097152 >> $O/check.log 2>&1
  echo "== $v C3 f32" >> $O/check.log
  timeout 300 python tools/gpu_check.py --config C3 --batch 32768 --f32 >> $O/check.log 2>&1
done
unset QLB_LIB
grep -E "==|rc=|device-resident|flag mism|grf rel|Error|error" $O/check.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=240 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qlb_single" -s 3 -c 1 -o $O/prof_main python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu_main.log 2>&1
ls $O
