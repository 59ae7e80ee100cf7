#!/usr/bin/env bash
# GPU call: validate the width-stacked tcgen05 convolution, compare back ends per layer, bench, ncu evidence.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_kernels.py -k "ws_tf32" -q --tb=short -p no:cacheprovider > $O/pytest_ws.log 2>&1
echo "pytest_ws rc=$?" >> $O/pytest_ws.log
tail -4 $O/pytest_ws.log
timeout 400 python tools/bench_conv.py all fp32,tc_tf32x3,ws_tf32x3,ws_tf32 > $O/bench_conv.log 2>&1
echo "bench_conv rc=$?" >> $O/bench_conv.log
timeout 400 python -m pytest tests/test_gpu_model.py -k "ws_tf32x3 or golden" -q --tb=short -p no:cacheprovider > $O/pytest_model_ws.log 2>&1
echo "pytest_model_ws rc=$?" >> $O/pytest_model_ws.log
tail -4 $O/pytest_model_ws.log
timeout 500 python bench.py --steps 10 --warmup 3 --no-alt-modes --dump-tuned $O/tuned.json > $O/bench.log 2>&1
echo "bench rc=$?" >> $O/bench.log
tail -2 $O/bench.log
DMVS_PRECISION=ws_tf32x3 timeout 300 python bench.py --steps 10 --warmup 3 --no-alt-modes --no-cpu-baseline > $O/bench_ws.log 2>&1
echo "bench_ws rc=$?" >> $O/bench_ws.log
tail -2 $O/bench_ws.log
# ncu: launch list of one step, then full captures of the dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-alt-modes --no-cpu-baseline > $O/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"get_cost|plane_sweep|depth_regression|upsample_depth|aggregate" -c 8 \
   -o $O/warp_full -f python bench.py --steps 1 --warmup 3 --no-alt-modes --no-cpu-baseline > $O/ncu_warp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_ws|conv_kernel|conv_tc" -c 12 \
   -o $O/conv_full -f env BENCH_CONV_REPS=1 python tools/bench_conv.py "feat.conv1.1,feat.out3,feat.conv0.1" fp32,ws_tf32x3 > $O/ncu_conv.log 2>&1
ls -la $O
