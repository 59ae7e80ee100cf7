#!/usr/bin/env bash
# GPU call: tests of the ws kernel, per-layer comparison, model parity, bench.
mkdir -p gpurun_out
O=gpurun_out
timeout 500 python -m pytest tests/test_gpu_kernels.py -k "ws_tf32 or auto or unshuffle or bn_folding" -q --tb=short -p no:cacheprovider > $O/pytest_ws.log 2>&1
echo "pytest_ws rc=$?" >> $O/pytest_ws.log
tail -4 $O/pytest_ws.log
timeout 300 python tools/bench_conv.py all fp32,tc_tf32x3,ws_tf32x3,ws_tf32 > $O/bench_conv.log 2>&1
echo "bench_conv rc=$?" >> $O/bench_conv.log
timeout 500 python -m pytest tests/test_gpu_model.py -k "ws_tf32x3 or golden or graph or feature" -q --tb=short -p no:cacheprovider > $O/pytest_model.log 2>&1
echo "pytest_model rc=$?" >> $O/pytest_model.log
tail -4 $O/pytest_model.log
timeout 500 python bench.py --steps 10 --warmup 3 --no-alt-modes --dump-tuned $O/tuned.json > $O/bench.log 2>&1
echo "bench rc=$?" >> $O/bench.log
tail -2 $O/bench.log | cut -c1-300
