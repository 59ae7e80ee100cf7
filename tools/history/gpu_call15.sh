#!/usr/bin/env bash
# PDL on every kernel of the step: full GPU suite + A/B
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
for pdl in 0 1 0 1; do
DMVS_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion --no-scan-mode > $O/bench_pdl$pdl.log 2>&1
grep '^{"metric' $O/bench_pdl$pdl.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pdl=$pdl:', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
