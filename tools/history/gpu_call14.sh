#!/usr/bin/env bash
# A/B: programmatic dependent launch on/off (DMVS_PDL), same box, same tuned table
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -x -k "cfg1 or graph or determin" > $O/pytest_pdl.log 2>&1
echo "pytest rc=$?" >> $O/pytest_pdl.log; tail -4 $O/pytest_pdl.log
for pdl in 0 1 0 1; do
DMVS_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion --no-scan-mode > $O/bench_pdl$pdl.log 2>&1
grep '^{"metric' $O/bench_pdl$pdl.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pdl=$pdl:', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
