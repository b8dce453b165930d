#!/bin/bash
# round-2 record run on one GPU: FP32-twin accuracy on C2 / C5, the whole GPU test suite, the bench, the launch list and
# one full ncu capture of the solve call on 2^20 C3 states
O=gpurun_out/exp17; mkdir -p $O
for c in C2 C5; do
  echo "== f32 $c" >> $O/f32.log
  timeout 300 python tools/gpu_check.py --config $c --batch 32768 --time-batch 1048576 --f32 >> $O/f32.log 2>&1
done
grep -E "==|grf rel|flag mism|margin|device-resident" $O/f32.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=240 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('value %.4g ms %.4f e2e %.4g (%s) other %.4g f32 %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['api'][:24], d['e2e']['other_entry']['value'], d['f32']['value']))"; tail -3 $O/bench.err
timeout 600 python bench.py --config C5 --steps 10 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; head -c 300 $O/bench_c5.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 > $O/b_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qlb_single|qlb_quad" -s 6 -c 2 -o $O/prof_full python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qlb_single|qlb_quad" -s 6 -c 2 -o $O/prof_full_f32 python tools/gpu_check.py --config C3 --batch 1024 --f32 > $O/ncu_full_f32.log 2>&1
ls $O
