#!/bin/bash
# usage: tools/r2_variants.sh OUTDIR "v1 v2 ..." [extra gpu_check args]   (main = the in-tree library)
O=gpurun_out/$1; mkdir -p $O; shift
VARS="$1"; shift
for v in $VARS; do
  if [ $v = main ]; then export QLB_LIB=$PWD/quadruped_locomotion_b200/libqlb.so; else export QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_$v.so; fi
  echo "== $v" >> $O/check.log
  timeout 300 python tools/gpu_check.py --config C3 --batch 8192 "$@" >> $O/check.log 2>&1; echo "rc=$?" >> $O/check.log
done
unset QLB_LIB
grep -E "==|device-resident|flag mism|Error|error" $O/check.log
