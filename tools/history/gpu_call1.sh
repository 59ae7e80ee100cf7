#!/usr/bin/env bash
# round 2, call 1: pipeline probe, full GPU parity suite (new operator-surface / cfg3 / cfg4 tests), cfg4 + cfg3 bench lines
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 300 ./tools/probes/pipe_probe > $O/pipe_probe.txt 2>&1
echo "probe rc=$?" >> $O/pipe_probe.txt
tail -60 $O/pipe_probe.txt
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -40 $O/pytest_gpu.log
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-alt-modes > $O/bench_cfg4.log 2>&1
echo "bench cfg4 rc=$?" >> $O/bench_cfg4.log
tail -2 $O/bench_cfg4.log | cut -c1-1500
timeout 600 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline > $O/bench_cfg3.log 2>&1
tail -1 $O/bench_cfg3.log | cut -c1-600
