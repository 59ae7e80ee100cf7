#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python tools/make_default_tuned.py 2>&1 | tail -3
cp diffmvs_b200/tuned/b200_default.json gpurun_out/b200_default.json
bash tools/gpu_round_check.sh
