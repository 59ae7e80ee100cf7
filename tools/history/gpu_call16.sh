#!/usr/bin/env bash
# experiment: upper bound of keeping the weight slabs resident (skip the per-stage weight fetch; results wrong)
mkdir -p gpurun_out
O=gpurun_out
DMVS_WS2_SKIPW=0 timeout 600 python tools/bench_conv.py all ws2_tf32x3 > $O/bench_conv_w1.txt 2>&1
DMVS_WS2_SKIPW=1 timeout 600 python tools/bench_conv.py all ws2_tf32x3 > $O/bench_conv_w0.txt 2>&1
paste <(cut -c1-44 $O/bench_conv_w1.txt) <(cut -c29-44 $O/bench_conv_w0.txt)
