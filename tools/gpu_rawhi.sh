#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
DMVS_WS_RAWHI=1 timeout 400 python -m pytest tests/test_gpu_kernels.py -k "ws_tf32x3" -q --tb=line -p no:cacheprovider > $O/pytest_rawhi.log 2>&1
tail -15 $O/pytest_rawhi.log
L="feat.conv0.1,feat.conv1.1,feat.out3,feat.conv2.1,feat.conv3.1,enc3 16->16,pvw 4->8"
timeout 200 python tools/bench_conv.py "$L" ws_tf32x3 > $O/rawhi0.log 2>&1
DMVS_WS_RAWHI=1 timeout 200 python tools/bench_conv.py "$L" ws_tf32x3 > $O/rawhi1.log 2>&1
paste $O/rawhi0.log $O/rawhi1.log | awk '{print $1,$2,$3,$4, $(NF-3)}' 
