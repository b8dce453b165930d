#!/bin/bash
# after the instruction-count work (no masking of inactive columns, normals staged by TMA, identity joint transforms
# skipped, frame in the core type for the FP32 twin, three CTAs per SM for FP32 arrays): parity + timing, tests, bench
O=gpurun_out/exp11; mkdir -p $O
for v in main f32c2; do
  if [ $v = main ]; then export QLB_LIB=$PWD/quadruped_locomotion_b200/libqlb.so; else export QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_$v.so; fi
  if [ $v = main ]; then
  echo "== $v C3" >> $O/check.log
  timeout 300 python tools/gpu_check.py --config C3 --batch 32768 >> $O/check.log 2>&1; echo "rc=$?" >> $O/check.log
  echo "== $v C5" >> $O/check.log
  timeout 300 python tools/gpu_check.py --config C5 --batch 32768 --time-batch 2097152 >> $O/check.log 2>&1
  fi
  echo "== $v C3 f32" >> $O/check.log
  timeout 300 python tools/gpu_check.py --config C3 --batch 32768 --f32 >> $O/check.log 2>&1
done
unset QLB_LIB
grep -E "==|rc=|device-resident|flag mism|grf rel|Error|error" $O/check.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=240 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('value %.4g ms %.4f e2e %.4g (%s) other %.4g f32 %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['api'][:24], d['e2e']['other_entry']['value'], d['f32']['value']))"; tail -3 $O/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qlb_single" -s 3 -c 1 -o $O/prof_main python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu_main.log 2>&1
ls $O
