#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -k "warp or plane_sweep or get_cost or golden or end_to_end or full_size" -q --tb=short -p no:cacheprovider > $O/pytest_warp.log 2>&1
echo "rc=$?" >> $O/pytest_warp.log
tail -5 $O/pytest_warp.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline > $O/bench_warp.log 2>&1
grep '^{' $O/bench_warp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['families_ms'])"
