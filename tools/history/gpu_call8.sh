#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -x -k "ws2" ) > $O/pytest_ws2.log 2>&1
echo "pytest rc=$?" >> $O/pytest_ws2.log
tail -5 $O/pytest_ws2.log
timeout 600 python tools/bench_conv.py all ws_tf32x3,ws2_tf32x3 > $O/bench_conv_ws2.txt 2>&1
cat $O/bench_conv_ws2.txt
DMVS_WS2_DBG=1 timeout 300 python tools/ws2_timeline.py "feat.conv1.1,feat.conv0.1" > $O/ws2_timeline.txt 2>&1
grep -E "^==|period" $O/ws2_timeline.txt | cut -c1-150
timeout 600 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --dump-tuned $O/tuned8.json > $O/bench_auto8.log 2>&1
grep '^{"metric' $O/bench_auto8.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('auto:', d['value'], d['ms_per_step'], d['e2e']['value'], d['scan_mode']['value'], d['roofline']['kernels_ms'])"
