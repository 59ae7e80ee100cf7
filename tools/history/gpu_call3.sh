#!/usr/bin/env bash
# ws2 kernel iteration: kernel parity tests for ws2, forced-ws2 bench, auto bench (tuning table), per-layer comparison
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -x -k "ws2" ) > $O/pytest_ws2.log 2>&1
echo "pytest rc=$?" >> $O/pytest_ws2.log
tail -5 $O/pytest_ws2.log
DMVS_PRECISION=ws2_tf32x3 timeout 600 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-scan-mode > $O/bench_ws2.log 2>&1
grep '^{"metric' $O/bench_ws2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('forced ws2:', d['value'], d['ms_per_step'], d['roofline']['kernels_ms'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --dump-tuned $O/tuned3.json > $O/bench_auto3.log 2>&1
grep '^{"metric' $O/bench_auto3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('auto:', d['value'], d['ms_per_step'], d['e2e']['value'], d['scan_mode'], d['roofline']['kernel'], d['roofline']['kernels_ms'])"
