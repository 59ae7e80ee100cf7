#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -x -k "up2 or fpn or paired" ) > $O/pytest_k5.log 2>&1
echo "pytest rc=$?" >> $O/pytest_k5.log; tail -25 $O/pytest_k5.log
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_operator_surface.py -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/pytest_k6.log 2>&1
echo "pytest rc=$?" >> $O/pytest_k6.log; tail -8 $O/pytest_k6.log
for c in 0 1 0 1; do
DMVS_FPN_COMPOSE=$c timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion --no-scan-mode > $O/bench_fpn$c.log 2>&1
grep '^{"metric' $O/bench_fpn$c.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('compose=$c:', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
