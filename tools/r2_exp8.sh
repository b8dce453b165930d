#!/bin/bash
O=gpurun_out/exp8; mkdir -p $O
timeout 120 python - > $O/small.log 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from quadruped_locomotion_b200 import capi, synth
s = capi.Solver("quadruped_model")
for B in (1, 2, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 66, 1003, 1004, 20001, 20002):
    ref = None
    for p in ("three_pass", "fused"):
        s.set_pipeline(p)
        out = s.solve_wrench_numpy(synth.make_states("C5", B, start=5))
        if ref is None: ref = out
        print(B, p, "ok", int(((out["flags"] >> 24) & 7 == 0).sum()), "max diff vs three_pass %.2e" % np.abs(out["grf"] - ref["grf"]).max(),
              "flags equal", bool(np.array_equal(out["flags"] & 0xFFFFFF, ref["flags"] & 0xFFFFFF)))
PY
echo "small rc=$?"; tail -2 $O/small.log
for v in main ctas2 ctas2s4; do
  if [ $v = main ]; then export QLB_LIB=$PWD/quadruped_locomotion_b200/libqlb.so; else export QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_$v.so; fi
  for p in fused fused_notma; do
    echo "== $v $p" | tee -a $O/check.log
    QLB_PIPELINE=$p timeout 300 python tools/gpu_check.py --config C3 --batch 32768 >> $O/check.log 2>&1; echo "rc=$?" >> $O/check.log
  done
  echo "== $v fused C5" >> $O/check.log
  timeout 300 python tools/gpu_check.py --config C5 --batch 32768 --time-batch 2097152 >> $O/check.log 2>&1
  echo "== $v fused C2" >> $O/check.log
  timeout 300 python tools/gpu_check.py --config C2 --batch 32768 --time-batch 65536 >> $O/check.log 2>&1
done
unset QLB_LIB
grep -E "==|rc=|device-resident|flag mism|Error|error" $O/check.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=240 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
for v in main ctas2; do
  if [ $v = main ]; then export QLB_LIB=$PWD/quadruped_locomotion_b200/libqlb.so; else export QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_$v.so; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlb_single -s 4 -c 1 -o $O/prof_$v python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu_$v.log 2>&1
done
unset QLB_LIB
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('value %.4g ms %.4f e2e %.4g (%s) other %.4g f32 %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['api'][:24], d['e2e']['other_entry']['value'], d['f32']['value']))"; tail -3 $O/bench.err
timeout 600 python bench.py --config C5 --steps 10 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; head -c 330 $O/bench_c5.json; tail -3 $O/bench_c5.err
ls $O
