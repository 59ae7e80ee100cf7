#!/usr/bin/env bash
# Final single-GPU evidence of the round: parity suite, smoke, default bench (both arms), cfg4 bench, parity report.
mkdir -p gpurun_out
O=gpurun_out
bash tools/gpu_round_check.sh
( time timeout 900 python bench.py --workload cfg4 --no-alt-modes --no-gpu-baseline ) > $O/bench_cfg4.log 2>&1
echo "cfg4 rc=$?"
grep '^{"metric' $O/bench_cfg4.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('cfg4 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
"
