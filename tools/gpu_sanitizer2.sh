#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
  python -m pytest tests/test_gpu_model.py tests/test_gpu_scan_pipeline.py -m gpu -q -x -p no:cacheprovider \
  -k "golden_reference or scene_loop_writes or scan_directory_to_point_cloud" > $O/sanitizer2.log 2>&1
echo "sanitizer rc=$?" >> $O/sanitizer2.log
grep -E 'ERROR SUMMARY|passed|failed|rc=|Invalid|out of bounds' $O/sanitizer2.log | head -20
