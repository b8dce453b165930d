#!/bin/bash
# round-2 experiment 1: baseline tests + bench, then kernel variants (timing + per-pass launch times)
mkdir -p gpurun_out/exp1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/exp1/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/exp1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp1/pytest.log
tail -3 gpurun_out/exp1/pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/exp1/bench.json 2> gpurun_out/exp1/bench.err; tail -c 600 gpurun_out/exp1/bench.json
for v in base ipm2 pdas8 pdas8ipm2; do
  if [ $v = base ]; then export QLB_LIB=$PWD/quadruped_locomotion_b200/libqlb.so; else export QLB_LIB=$PWD/quadruped_locomotion_b200/variants/libqlb_$v.so; fi
  echo "== $v" | tee -a gpurun_out/exp1/variants.log
  python tools/gpu_check.py --config C3 --batch 32768 >> gpurun_out/exp1/variants.log 2>&1
  python tools/gpu_check.py --config C5 --batch 32768 --time-batch 2097152 >> gpurun_out/exp1/variants.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qlb_quad -s 9 -c 6 --csv --log-file gpurun_out/exp1/launches_$v.csv python tools/gpu_check.py --config C3 --batch 1024 > /dev/null 2>&1
  grep -E "qlb_quad" gpurun_out/exp1/launches_$v.csv | awk -F'","' '{print $5, $NF}' | tail -6 | tee -a gpurun_out/exp1/variants.log
done
grep -E "==|device-resident|status hist|flag mism|grf rel" gpurun_out/exp1/variants.log
