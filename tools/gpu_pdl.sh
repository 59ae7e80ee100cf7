#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
DMVS_PDL=1 timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -k "ws_tf32x3 or auto or graph or golden" -q --tb=line -p no:cacheprovider > $O/pytest_pdl.log 2>&1
tail -4 $O/pytest_pdl.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline --dump-tuned $O/tuned.json > $O/bench_pdl0.log 2>&1
DMVS_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline --load-tuned $O/tuned.json > $O/bench_pdl1.log 2>&1
DMVS_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline --load-tuned $O/tuned.json --no-graph > $O/bench_pdl1_nograph.log 2>&1
for f in bench_pdl0 bench_pdl1 bench_pdl1_nograph; do echo $f; grep '^{' $O/$f.log | cut -c1-330; done
