#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > $O/bench_2gpu.log 2>&1
echo "rc=$?" >> $O/bench_2gpu.log
grep '^{"metric' $O/bench_2gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'n', d['n_gpus'])
print('fusion', d['fusion'])
"
tail -5 $O/bench_2gpu.log | cut -c1-300
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > $O/bench_ref_2gpu.log 2>&1
tail -3 $O/bench_ref_2gpu.log | cut -c1-200
