#!/usr/bin/env bash
DMVS_TIMELINE_MODE=ws2_f16c DMVS_WS2_DBG=1 timeout 300 python tools/ws2_timeline.py "feat.conv1.1,feat.out2,unet3.init" 2>&1 | cut -c1-190
