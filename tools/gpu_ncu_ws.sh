#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_ws" -c 4 \
   -o $O/ws_full -f env BENCH_CONV_REPS=1 python tools/bench_conv.py "feat.out3,feat.conv1.1" ws_tf32x3 > $O/ncu_ws.log 2>&1
ls -la $O/ws_full.ncu-rep
