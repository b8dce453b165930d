#!/bin/bash
# record run on one GPU: timing, the whole GPU test suite, bench, launch list, full ncu captures (FP64 and FP32 twin), all configs
O=gpurun_out/${1:-exp70}; mkdir -p $O
bash tools/r2_variants.sh ${1:-exp70} "main main"
bash tools/r2_variants.sh ${1:-exp70}f "main" --f32
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=240 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 400 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('value %.4g ms %.4f e2e %.4g other %.4g f32 %.4g frac %.3f exec %s traffic %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['other_entry']['value'], d['f32']['value'], d['roofline']['frac'], d['roofline']['executed']['fp64_pipe_active_pct_of_elapsed'], d['roofline']['traffic']))"; tail -3 $O/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 > $O/b_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"qlb_single|qlb_quad" -s 6 -c 2 -o $O/prof_full python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"qlb_single|qlb_quad" -s 6 -c 2 -o $O/prof_full_f32 python tools/gpu_check.py --config C3 --batch 1024 --f32 > $O/ncu_full_f32.log 2>&1
timeout 200 python tools/configs_report.py > $O/configs.jsonl 2> $O/configs.err; cat $O/configs.jsonl | cut -c1-300
ls $O
