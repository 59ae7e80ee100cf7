#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -x -k "paired or conv3d_matches" ) 2>&1 | tail -12
timeout 300 python tools/bench_conv.py "feat.conv0.0,pvw 4" ws2_tf32x3,ws2_f16c 2>&1 | tail -4
