#!/usr/bin/env bash
# run a subset of the GPU tests: tools/gpu_tests_some.sh <pytest args>
mkdir -p gpurun_out
( timeout 1200 python -m pytest "$@" -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest_some.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_some.log
tail -40 gpurun_out/pytest_some.log
