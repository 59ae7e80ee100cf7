#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -x -k "ws2_f16c" ) > $O/pytest_f16c.log 2>&1
echo "pytest rc=$?" >> $O/pytest_f16c.log; tail -25 $O/pytest_f16c.log
timeout 600 python tools/bench_conv.py all ws2_tf32x3,ws2_f16c 2>&1 | tail -32
