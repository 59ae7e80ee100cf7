#!/usr/bin/env bash
mkdir -p gpurun_out
DMVS_WS2_DBG=1 timeout 300 python tools/ws2_timeline.py "feat.conv1.1,feat.conv0.1,feat.inner2" > gpurun_out/ws2_timeline.txt 2>&1
cat gpurun_out/ws2_timeline.txt
