#!/bin/bash
# which of the two setup changes costs time: unit-vector renormalisation (vC), trivial-transform skip (vB), both (main), neither (vA)
O=gpurun_out/exp12; mkdir -p $O
for v in vA main vB vC vA main; do
  if [ $v = main ]; then export QLB_LIB=$PWD/quadruped_locomotion_b200/libqlb.so; else export QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_$v.so; fi
  echo "== $v C3" >> $O/check.log
  timeout 300 python tools/gpu_check.py --config C3 --batch 8192 >> $O/check.log 2>&1; echo "rc=$?" >> $O/check.log
done
unset QLB_LIB
grep -E "==|device-resident|flag mism|grf rel|Error|error" $O/check.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=240 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
