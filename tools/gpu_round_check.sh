#!/usr/bin/env bash
# Final check of the round: full GPU suite, smoke, bench (both arms), full ncu capture of the ws kernel in the model's shapes.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/smoke.log
tail -2 $O/smoke.log
timeout 600 python bench.py --dump-tuned $O/tuned.json > $O/bench.log 2>&1
echo "bench rc=$?" >> $O/bench.log
tail -2 $O/bench.log | cut -c1-300
true
ls -la $O | head -20
timeout 200 python tools/bench_fusion.py > $O/fusion.log 2>&1
tail -1 $O/fusion.log | cut -c1-600
