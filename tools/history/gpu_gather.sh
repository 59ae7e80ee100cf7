#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -x -k "get_cost or plane_sweep or warp" ) 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion --no-scan-mode --no-batched > $O/bench_gather.log 2>&1
grep '^{"metric' $O/bench_gather.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench:', d['value'], d['ms_per_step']); k=d['roofline']['kernels_ms']; print({n:k[n] for n in ('get_cost','plane_sweep_corr')})"
