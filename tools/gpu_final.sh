#!/usr/bin/env bash
# Round-end style check: full GPU test suite, smoke, both bench arms, ncu launch list and DRAM traffic of one step.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/smoke.log
tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench.log 2>&1
echo "bench rc=$?" >> $O/bench.log
tail -2 $O/bench.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.log 2>&1
echo "bench_ref rc=$?" >> $O/bench_ref.log
tail -2 $O/bench_ref.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-alt-modes --no-cpu-baseline --no-graph > $O/ncu_list.log 2>&1
echo "ncu rc=$?" >> $O/ncu_list.log
ls -la $O | head -30
