#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/pytest_kernels.log 2>&1
echo "pytest rc=$?" >> $O/pytest_kernels.log; tail -15 $O/pytest_kernels.log
timeout 300 python tools/bench_conv.py "feat.conv0.0,feat.conv0.1,pvw,reg 8" ws_tf32x3,ws2_tf32x3,fp32
DMVS_WS2_DBG=1 timeout 300 python tools/ws2_timeline.py "feat.conv0.0,pvw 4->8" 2>&1 | cut -c1-200
