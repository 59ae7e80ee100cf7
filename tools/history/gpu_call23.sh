#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_operator_surface.py -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/pytest_k7.log 2>&1
echo "pytest rc=$?" >> $O/pytest_k7.log; tail -8 $O/pytest_k7.log
for c in 0 1 0 1; do
DMVS_UNET_UP2=$c timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion --no-scan-mode > $O/bench_up2$c.log 2>&1
grep '^{"metric' $O/bench_up2$c.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('unet_up2=$c:', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
