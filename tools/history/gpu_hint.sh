#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/bench_conv.py "feat.conv0.1,feat.conv1.1,feat.conv2.1,feat.conv3.1,feat.out2,unet3.init,unet 32,enc2 32,pvw 4,gru.zr" auto 2>&1 | tail -12
timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion --no-scan-mode --no-batched > gpurun_out/bench_hint.log 2>&1
grep '^{"metric' gpurun_out/bench_hint.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench:', d['value'], d['ms_per_step'])"
