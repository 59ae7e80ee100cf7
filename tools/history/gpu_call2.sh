#!/usr/bin/env bash
# round 2, call 2: MMA chain probe, ws2 kernel tests, benches with ws2 forced / auto
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./tools/probes/pipe_probe > $O/pipe_probe2.txt 2>&1
echo "probe rc=$?" >> $O/pipe_probe2.txt
grep -E "chain|^tma" $O/pipe_probe2.txt | head -70
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_scan_fusion.py tests/test_gpu_fusion.py -m gpu -q --tb=short -p no:cacheprovider -x -k "ws2 or scan or fus or cache or uint8 or noise" ) > $O/pytest_ws2.log 2>&1
echo "pytest rc=$?" >> $O/pytest_ws2.log
tail -30 $O/pytest_ws2.log
DMVS_PRECISION=ws2_tf32x3 timeout 600 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-scan-mode > $O/bench_ws2.log 2>&1
tail -1 $O/bench_ws2.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline --dump-tuned $O/tuned2.json > $O/bench_auto2.log 2>&1
tail -1 $O/bench_auto2.log | cut -c1-3000
