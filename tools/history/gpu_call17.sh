#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
DMVS_WS2_DBG=1 timeout 300 python tools/ws2_timeline.py "feat.conv0.0,feat.conv0.1,pvw 4->8,pvw 8->1,feat.out3,feat.conv2.1" > $O/ws2_timeline_big.txt 2>&1
cut -c1-200 $O/ws2_timeline_big.txt
timeout 300 python tools/bench_conv.py "feat.conv0.0,feat.conv0.1" ws_tf32x3,ws2_tf32x3,fp32
