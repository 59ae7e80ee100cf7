#!/usr/bin/env bash
# Round-2 ncu --set full captures: the TMA-fed convolution on four representative layers, and the non-convolution
# kernels of the path from one eager cfg3 step.  Reports land in gpurun_out/ (summarised into profiles/ afterwards).
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_ws2" -c 4 \
   -o $O/r2_ws2_final -f env BENCH_CONV_REPS=1 python tools/bench_conv.py "feat.conv1.1,feat.out3,feat.conv0.0,pvw 4->8" auto > $O/ncu_ws2_final.log 2>&1
echo "ncu ws2 rc=$?"
timeout 400 python bench.py --steps 3 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-scan-mode --no-fusion --no-batched --dump-tuned $O/tuned.json > $O/bench_short.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"get_cost|plane_sweep|conv3d_to1|deconv3d_parity|groupnorm_silu" -c 9 \
   -o $O/r2_other_final -f python bench.py --steps 1 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-scan-mode --no-fusion --no-batched --no-graph --load-tuned $O/tuned.json > $O/ncu_other_final.log 2>&1
echo "ncu other rc=$?"
ls -la $O/*.ncu-rep
