#!/bin/bash
# multi-GPU record run (run with gpurun --gpus N): C3 bench and the C5 sweep under torchrun on all N GPUs, the C++ NCCL
# statistics demo, and the pinned-copy bandwidth with 1, 2 and N ranks copying at once (the end-to-end limiter)
N=${1:-8}
O=gpurun_out/exp43_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 240 $TR bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
python -c "
import json; d=json.load(open('$O/bench_n$N.json')); print('C3 n=%d value %.4g ms %.4f e2e %.4g f32 %.4g' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['f32']['value']))"
timeout 240 $TR bench.py --config C5 --gpus $N --steps 10 --warmup 3 > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err
python -c "
import json; d=json.load(open('$O/bench_c5_n$N.json')); print('C5 n=%d value %.4g ms %.4f e2e %.4g states %d' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['global_batch']))"
timeout 120 python tools/pcie_check.py > $O/pcie_1.json 2> $O/pcie.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/pcie_check.py > $O/pcie_2.json 2>> $O/pcie.err
if [ $N -gt 2 ]; then timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/pcie_check.py > $O/pcie_$N.json 2>> $O/pcie.err; fi
cat $O/pcie_*.json | cut -c1-400
python -c "
from quadruped_locomotion_b200 import build; print(build.build_host_demo(which='nccl_demo'))" > /dev/null 2>&1
timeout 300 quadruped_locomotion_b200/host/nccl_demo 1048576 > $O/nccl_demo.txt 2>&1; tail -1 $O/nccl_demo.txt
ls $O
