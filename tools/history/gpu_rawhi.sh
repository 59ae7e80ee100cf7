#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/pytest_rawhi.log 2>&1
echo "pytest rc=$?" >> $O/pytest_rawhi.log; tail -6 $O/pytest_rawhi.log
grep -E 'ws2_f16c|cfg3|cfg4' $O/parity_report.txt | head -40
timeout 600 python tools/bench_conv.py "feat.conv0.1,feat.conv1.1,feat.conv2.1,feat.conv3.1,feat.out2,unet3.init,unet 32,enc2 32,gru.zr" ws2_f16c 2>&1 | tail -10
timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion --no-scan-mode --no-batched > $O/bench_rawhi.log 2>&1
grep '^{"metric' $O/bench_rawhi.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench:', d['value'], d['ms_per_step'])"
