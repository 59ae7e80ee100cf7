#!/usr/bin/env bash
# bench lines of the other named workloads (cfg2: DiffMVS 640x512x5 views, cfg4: CasDiffMVS 1920x1024x11 views)
mkdir -p gpurun_out
O=gpurun_out
for wl in cfg2 cfg4; do
( time timeout 900 python bench.py --workload $wl --no-alt-modes ) > $O/bench_$wl.log 2>&1
echo "$wl rc=$?"
grep '^{"metric' $O/bench_$wl.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'gpu_baseline', d['gpu_baseline'] and d['gpu_baseline'].get('tf32_default'), 'batched', d['batched'] and d['batched']['value'])
"
done
