#!/usr/bin/env bash
# Profiles of the final state: tuned table from a normal run, ncu launch list of one eager step with that table, full
# captures of the dominant convolution kernels.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 --no-alt-modes --no-cpu-baseline --dump-tuned $O/tuned.json > $O/bench_short.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-alt-modes --no-cpu-baseline --no-graph --load-tuned $O/tuned.json > $O/ncu_list.log 2>&1
echo "ncu list rc=$?" >> $O/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_ws|conv_tc|conv_kernel" -c 8 \
   -o $O/conv_final -f env BENCH_CONV_REPS=1 python tools/bench_conv.py "feat.conv1.1,feat.out3,feat.conv3.1,feat.inner2" auto > $O/ncu_conv.log 2>&1
ls -la $O | head
