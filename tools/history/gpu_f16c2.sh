#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -k "tensor_core_modes or zz_report" ) > $O/pytest_f16c2.log 2>&1
echo "pytest rc=$?" >> $O/pytest_f16c2.log; tail -5 $O/pytest_f16c2.log
grep -E 'ws2_tf32x3|ws2_f16c|\[fp32\]' $O/parity_report.txt | head -60
for m in ws2_tf32x3 ws2_f16c ws2_tf32x3 ws2_f16c; do
DMVS_PRECISION=$m timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion --no-scan-mode --no-batched > $O/bench_$m.log 2>&1
grep '^{"metric' $O/bench_$m.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m:', d['value'], d['ms_per_step'])"
done
