#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ws2 -c 4 -o $O/ws2_full2 -f \
   env BENCH_CONV_REPS=1 python tools/bench_conv.py "feat.conv1.1,feat.conv0.1" ws2_tf32x3 > $O/ncu_ws2.log 2>&1
tail -3 $O/ncu_ws2.log
