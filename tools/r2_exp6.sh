#!/bin/bash
O=gpurun_out/exp6; mkdir -p $O
# small / ragged batches first, bounded: this is where the previous build hung
timeout 120 python - > $O/small.log 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from quadruped_locomotion_b200 import capi, synth
s = capi.Solver("quadruped_model")
for B in (1, 2, 7, 8, 9, 15, 16, 17, 63, 64, 65, 1003):
    for p in ("fused", "three_pass"):
        s.set_pipeline(p)
        out = s.solve_wrench_numpy(synth.make_states("C5", B, start=5))
        print(B, p, "ok", int(((out["flags"] >> 24) & 7 == 0).sum()))
PY
echo "small rc=$?"; tail -3 $O/small.log
for v in main super1; do
  if [ $v = main ]; then export QLB_LIB=$PWD/quadruped_locomotion_b200/libqlb.so; else export QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_$v.so; fi
  for p in fused fused_notma; do
    echo "== $v $p" | tee -a $O/check.log
    QLB_PIPELINE=$p timeout 300 python tools/gpu_check.py --config C3 --batch 32768 >> $O/check.log 2>&1; echo "rc=$?" >> $O/check.log
  done
done
export QLB_LIB=$PWD/quadruped_locomotion_b200/libqlb.so
echo "== main three_pass" >> $O/check.log
QLB_PIPELINE=three_pass timeout 300 python tools/gpu_check.py --config C3 --batch 32768 >> $O/check.log 2>&1
echo "== main fused C5" >> $O/check.log
timeout 300 python tools/gpu_check.py --config C5 --batch 32768 --time-batch 2097152 >> $O/check.log 2>&1
echo "== main fused C2" >> $O/check.log
timeout 300 python tools/gpu_check.py --config C2 --batch 32768 --time-batch 65536 >> $O/check.log 2>&1
grep -E "==|rc=|device-resident|flag mism|Error|error" $O/check.log
unset QLB_LIB
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=240 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
for p in fused fused_notma; do
QLB_PIPELINE=$p timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlb_single -s 4 -c 1 -o $O/prof_$p python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu_$p.log 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; head -c 300 $O/bench.json; tail -3 $O/bench.err
ls -la $O
