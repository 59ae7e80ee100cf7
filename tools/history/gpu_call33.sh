#!/usr/bin/env bash
# experiment: one MMA-issuing warp with / without collector reuse of the A_hi operand vs four issuing warps
mkdir -p gpurun_out
O=gpurun_out
L="feat.conv0.1,feat.conv1.1,feat.conv2.1,feat.conv3.1,feat.out3,feat.out2,unet 32,enc2 32,pvw 4"
DMVS_WS2_MMAW=4 timeout 600 python tools/bench_conv.py "$L" ws2_tf32x3 > $O/bc_w4.txt 2>&1
DMVS_WS2_MMAW=1 DMVS_WS2_AREUSE=0 timeout 600 python tools/bench_conv.py "$L" ws2_tf32x3 > $O/bc_w1.txt 2>&1
DMVS_WS2_MMAW=1 DMVS_WS2_AREUSE=1 timeout 600 python tools/bench_conv.py "$L" ws2_tf32x3 > $O/bc_w1r.txt 2>&1
paste <(cut -c1-44 $O/bc_w4.txt) <(cut -c29-44 $O/bc_w1.txt) <(cut -c29-44 $O/bc_w1r.txt)
DMVS_WS2_MMAW=1 DMVS_WS2_AREUSE=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
