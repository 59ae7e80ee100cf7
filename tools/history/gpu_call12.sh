#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_operator_surface.py -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/pytest_model.log 2>&1
echo "pytest rc=$?" >> $O/pytest_model.log
tail -6 $O/pytest_model.log
for br in 0 1; do
DMVS_BRANCHES=$br timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion > $O/bench_br$br.log 2>&1
grep '^{"metric' $O/bench_br$br.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('branches=$br:', d['value'], d['ms_per_step'], d['e2e']['value'], d['scan_mode']['value'])"
done
