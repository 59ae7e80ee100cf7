#!/usr/bin/env bash
# launch list + ncu --set full captures of the final tree
bash tools/gpu_launchlist.sh
bash tools/gpu_profiles_r2.sh
