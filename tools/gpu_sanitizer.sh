#!/usr/bin/env bash
# compute-sanitizer memcheck over the kernels added this round (small test cases)
mkdir -p gpurun_out
O=gpurun_out
DMVS_AUTOTUNE=0 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider \
  -k "conv3d_to1 or deconv3d or paired or up2 or fpn or uint8 or (conv2d_matches_torch and ws2_f16c and 16-16)" > $O/sanitizer.log 2>&1
echo "sanitizer rc=$?" >> $O/sanitizer.log
grep -E 'ERROR SUMMARY|passed|failed|rc=|Invalid|out of bounds' $O/sanitizer.log | head -20
