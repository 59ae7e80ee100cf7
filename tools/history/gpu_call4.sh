#!/usr/bin/env bash
# ncu of the ws2 kernel on three representative layers + layer-by-layer micro-benchmark ws vs ws2
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/bench_conv.py all ws_tf32x3,ws2_tf32x3 > $O/bench_conv_ws2.txt 2>&1
cat $O/bench_conv_ws2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ws2 -c 6 -o $O/ws2_full -f \
   env BENCH_CONV_REPS=1 python tools/bench_conv.py "feat.conv1.1,feat.conv0.1,feat.out3" ws2_tf32x3 > $O/ncu_ws2.log 2>&1
tail -3 $O/ncu_ws2.log
ls -la $O/*.ncu-rep
