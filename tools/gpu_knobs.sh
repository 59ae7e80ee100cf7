#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
L="feat.conv0.1,feat.conv1.1,feat.out3,feat.conv2.1,unet3.rb 8->8,enc3 16->16,pvw 4->8"
i=0
for e in "X=1" "DMVS_WS_R=3" "DMVS_WS_TH=4" "DMVS_WS_TH=4 DMVS_WS_R=3" "DMVS_WS_TH=2 DMVS_WS_R=3" "DMVS_WS_TH=16"; do
  echo "== $e" > $O/knob_$i.log
  env $e timeout 200 python tools/bench_conv.py "$L" ws_tf32x3,ws_tf32 >> $O/knob_$i.log 2>&1
  i=$((i+1))
done
cat $O/knob_*.log
