#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
L="feat.conv0.1,feat.conv1.0,feat.conv1.1,feat.out3,feat.conv2.1,feat.conv3.1,unet3.rb 8->8,enc3 16->16,pvw 4->8,unet 32->32"
i=0
for e in "X=1" "DMVS_WS_ALIGN8=1" "DMVS_WS_CTAS=1" "DMVS_WS_CTAS=1 DMVS_WS_ALIGN8=1"; do
  echo "== $e" > $O/knob_$i.log
  env $e timeout 200 python tools/bench_conv.py "$L" ws_tf32x3 >> $O/knob_$i.log 2>&1
  i=$((i+1))
done
cat $O/knob_0.log $O/knob_1.log $O/knob_2.log $O/knob_3.log
