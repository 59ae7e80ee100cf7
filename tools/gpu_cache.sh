#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_model.py -k "feature_cache or graph or golden" -q --tb=short -p no:cacheprovider > $O/pytest_cache.log 2>&1
echo "rc=$?" >> $O/pytest_cache.log
tail -6 $O/pytest_cache.log
timeout 300 python tools/bench_scan_cache.py cfg3 10 > $O/scan_cache.log 2>&1
tail -2 $O/scan_cache.log
