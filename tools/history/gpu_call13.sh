#!/usr/bin/env bash
mkdir -p gpurun_out
DMVS_WS2_DBG=1 timeout 300 python tools/ws2_timeline.py "unet 32->32 @1/8,gru.zr,unet3.rb 8->8,enc2 32->32" > gpurun_out/ws2_timeline_small.txt 2>&1
cat gpurun_out/ws2_timeline_small.txt | cut -c1-230
