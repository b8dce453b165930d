#!/bin/bash
# full GPU test suite + bench + launch list + one full ncu capture of the solve call on 2^20 C3 states (default library)
O=gpurun_out/exp9; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=240 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('value %.4g ms %.4f e2e %.4g (%s) other %.4g f32 %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['api'][:24], d['e2e']['other_entry']['value'], d['f32']['value']))"; tail -3 $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 > $O/b_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qlb_single|qlb_quad" -s 6 -c 2 -o $O/prof_full python tools/gpu_check.py --config C3 --batch 1024 > $O/ncu_full.log 2>&1
ls $O
