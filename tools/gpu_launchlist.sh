#!/usr/bin/env bash
# ncu launch list (duration + DRAM bytes per launch) of one eager cfg3 step run with the back ends a normal run chose
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-scan-mode --no-fusion --no-batched --dump-tuned $O/tuned.json > $O/bench_short.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-scan-mode --no-fusion --no-batched --no-graph --load-tuned $O/tuned.json > $O/ncu_list.log 2>&1
echo "ncu list rc=$?" >> $O/ncu_list.log
tail -2 $O/ncu_list.log | cut -c1-200
wc -l $O/launches.csv
