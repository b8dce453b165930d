N=${1:-2}; O=gpurun_out/exp71_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
python -c "
import json; d=json.load(open('$O/bench_n$N.json')); print('C3 n=%d value %.4g ms %.4f e2e %.4g f32 %.4g' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['f32']['value']))"
timeout 240 $TR bench.py --config C5 --gpus $N --steps 10 --warmup 3 > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err
python -c "
import json; d=json.load(open('$O/bench_c5_n$N.json')); print('C5 n=%d value %.4g ms %.4f e2e %.4g states %d' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['global_batch']))"
